"""CPU tests of the host side: module surface, update coefficients, C-ABI exports, sharding (gloo)."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from diffroll_b200.synthetic import default_hparams, make_inputs, make_state_dict, state_dict_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(**kw):
    import diffroll_b200 as M
    hp = default_hparams(**kw)
    return M.ClassifierFreeDiffRoll(**hp), hp


# ---- module surface ------------------------------------------------------------------------------------
def test_class_lookup_like_sampling_py():
    import diffroll_b200 as Model
    assert getattr(Model, "ClassifierFreeDiffRoll").__name__ == "ClassifierFreeDiffRoll"   # sampling.py:54


def test_state_dict_keys_and_shapes_match_reference_contract():
    m, hp = _model()
    sd = m.state_dict()
    want = state_dict_shapes(hp)
    assert len(sd) == 132 and set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert "diffusion_embedding.embedding" not in sd            # persistent=False, model/diffwave.py:61
    assert float(m.output_projection.weight.abs().max()) == 0.0  # zero-initialised head, model/diffwave.py:630
    m.load_state_dict(make_state_dict(hp), strict=True)


def test_hparams_and_errors():
    m, hp = _model(inpainting_t=[0, 320])
    assert m.hparams.timesteps == 200 and m.hparams.sampling.type == "inpainting_ddpm_x0"
    assert m.hparams.sampling.w == 0.5 and m.hparams.inpainting_t == [0, 320] and m.hparams.condition == "fixed"
    assert m.hparams.spec_args.n_mels == 229
    import diffroll_b200 as M
    bad = default_hparams(); bad["condition"] = "nonsense"
    with pytest.raises(ValueError):
        M.ClassifierFreeDiffRoll(**bad)
    bad = default_hparams(sampling_type="no_such_sampler")
    with pytest.raises(AttributeError):
        M.ClassifierFreeDiffRoll(**bad)
    with pytest.raises(NotImplementedError):
        m.p_losses(torch.zeros(1), torch.zeros(1), loss_type="nope")


def test_load_from_checkpoint_overrides(tmp_path):
    import diffroll_b200 as M
    hp = default_hparams(residual_layers=2)
    sd = make_state_dict(hp)
    path = tmp_path / "tiny.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": hp}, path)
    m = M.ClassifierFreeDiffRoll.load_from_checkpoint(str(path), sampling={"type": "generation_ddpm_x0"},
                                                      inpainting_t=None, generation_filter=0.1)
    assert m.hparams.sampling.type == "generation_ddpm_x0" and m.hparams.generation_filter == 0.1
    assert m.reverse_diffusion.__name__ == "generation_ddpm_x0"
    assert torch.equal(m.state_dict()["skip_projection.weight"], sd["skip_projection.weight"])


def test_schedule_and_embedding_tables_equal_oracle():
    from oracle.diffroll_oracle import Schedule, build_embedding
    for T in (200, 1000):
        m, hp = _model(timesteps=T)
        s = Schedule(hp["beta_start"], hp["beta_end"], T)
        for name in ("betas", "sqrt_recip_alphas", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "posterior_variance", "alphas"):
            assert torch.equal(getattr(m, name), getattr(s, name)), name
        assert torch.equal(m.diffusion_embedding.embedding, build_embedding(T))


# ---- update coefficients: the fused epilogue formula reproduces the oracle's sampler arithmetic ------------
def _apply(u, net, x, n):
    s = [np.float32(v) for v in u.s]
    net, x, n = net.astype(np.float32), x.astype(np.float32), n.astype(np.float32)
    if u.mode == 0:
        r = s[0] * net + s[1] * (x - s[2] * net) / s[3]
        return r + s[4] * n if u.has_noise else r
    if u.mode == 1:
        return net / s[0]
    if u.mode == 2:
        r = s[0] * (x - s[1] * net / s[2])
        return r + s[3] * n if u.has_noise else r
    if u.mode == 3:
        r = s[0] * ((x - s[1] * net) / s[2]) + s[3] * net
        return r + s[4] * n if u.has_noise else r
    if u.mode == 4:
        return (x - s[0] * net) / s[1]
    return net


@pytest.mark.parametrize("name", ["inpainting_ddpm_x0", "cfdg_ddpm_x0", "generation_ddpm_x0", "ddpm_x0", "ddim_x0",
                                  "cfdg_ddim_x0", "ddpm", "ddim", "ddim2ddpm"])
def test_update_structs_reproduce_oracle_samplers(name):
    from diffroll_b200 import _lib
    from oracle.diffroll_oracle import OracleDiffRoll
    m, hp = _model(sampling_type=name)
    ups, branches, masks = m._all_updates()
    assert len(ups) == 200
    want_branch = {"inpainting_ddpm_x0": _lib.BRANCH_COND_UNCOND, "cfdg_ddpm_x0": _lib.BRANCH_COND_UNCOND,
                   "generation_ddpm_x0": _lib.BRANCH_UNCOND, "cfdg_ddim_x0": _lib.BRANCH_COND_ZEROSPEC}.get(name, _lib.BRANCH_COND)
    assert branches == want_branch

    class Fake(OracleDiffRoll):          # network output replaced by a fixed tensor: isolates the posterior arithmetic
        def forward(self, x_t, waveform, diffusion_step, **kw):
            return self.net, None
        __call__ = forward

    o = Fake(hp, {k: v for k, v in make_state_dict(hp).items() if k.startswith("mel")})
    g = torch.Generator().manual_seed(0)
    net, x, n = (torch.randn(1, 1, 8, 88, generator=g) for _ in range(3))
    o.net = net
    for t_index in (199, 100, 1, 0):
        u = ups[199 - t_index]
        ref, _ = o.reverse_diffusion(x, torch.zeros(1, 4), t_index, noise=n)
        got = _apply(u, net.numpy(), x.numpy(), n.numpy())
        assert np.abs(got - ref.numpy()).max() < 2e-6 * max(1.0, float(ref.abs().max())), (name, t_index)
        assert bool(u.has_noise) == (t_index > 0 and name not in ("ddim_x0", "cfdg_ddim_x0", "ddim"))


def test_update_structs_are_memoised_and_follow_what_the_samplers_read():
    """sample_loop asks for the 200 update structs at the head of every call (10 ms of host arithmetic): they are computed once and
    recomputed when anything a sampler method reads changes -- guidance weight, sampler, mask ranges, the schedule tensors."""
    m, hp = _model(sampling_type="inpainting_ddpm_x0")
    ups, branches, masks = m._all_updates()
    assert m._all_updates()[0] is ups                       # memo hit: the same list object
    w0 = ups[0].w
    m.hparams.sampling.w = w0 + 1.5
    ups2, _, _ = m._all_updates()
    assert ups2 is not ups and abs(ups2[0].w - (w0 + 1.5)) < 1e-6
    m.hparams.inpainting_t = [0, 320]
    assert m._all_updates()[2][0] == [0, 320]
    m.betas.mul_(1.0)                                       # an in-place edit of a schedule tensor bumps its version
    ups3, _, _ = m._all_updates()
    assert ups3 is not ups2
    m.sqrt_alphas_cumprod = m.sqrt_alphas_cumprod * 0.5     # a replaced schedule tensor
    ups4, _, _ = m._all_updates()
    assert abs(ups4[-1].s[0] - 0.5 * ups3[-1].s[0]) < 1e-7  # t = 0: x_0 scaled by sqrt_alphas_cumprod[0]


def test_no_cpu_fallback_and_training_mode_guard():
    from diffroll_b200._lib import DrbError
    m, hp = _model()
    x, w = torch.randn(1, 1, 128, 88), torch.randn(1, 65536)
    with pytest.raises(DrbError):
        m(x, w, torch.tensor([3]))                 # training mode (spec dropout + saved activations) is CUDA-only as well
    with pytest.raises(DrbError):
        m.training_step({"frame": torch.zeros(1, 128, 88), "audio": w}, 0)
    m.eval()
    with pytest.raises(DrbError):
        m(x, w, torch.tensor([3]))                 # CPU tensors: no fallback
    with pytest.raises(DrbError):
        m.reverse_diffusion(x, w, 5)


def test_unconditional_blocks_behave_like_the_reference():
    """unconditional=True: the reference builds blocks without conditioner_projection (122 state_dict tensors instead of 132 + 2
    buffers... ) and its forward then fails ResidualBlock's assertion (model/diffwave.py:135-136) because ClassifierFreeDiffRoll
    always passes a spectrogram.  Same construction, same failure."""
    import diffroll_b200 as M
    hp = default_hparams()
    hp["unconditional"] = True
    m = M.ClassifierFreeDiffRoll(**hp)
    keys = set(m.state_dict())
    assert not any("conditioner_projection" in k for k in keys) and len(keys) == 132 - 2 * hp["residual_layers"]
    with pytest.raises(AssertionError):
        m.eval()(torch.randn(1, 1, 128, 88), torch.randn(1, 65536), torch.tensor([3]))


# ---- C ABI ------------------------------------------------------------------------------------------------------
def test_cabi_exports_every_declared_symbol(lib_built):
    header = open(os.path.join(ROOT, "include", "diffroll_b200.h")).read()
    declared = set(re.findall(r"\b(drb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    lib = C.CDLL(lib_built)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/diffroll_b200.h but not exported"
    from diffroll_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().drb_version() == 100


def test_cabi_workspace_query_and_argument_checks(lib_built):
    from diffroll_b200 import _lib
    lib = _lib.load()
    cfg = _lib.DrbConfig(batch=32, frames=640, pitches=88, wave_len=327680, residual_channels=512, residual_layers=15,
                         kernel_size=9, dilation_base=2, dilation_bound=4, n_mels=229, n_fft=2048, hop_length=512,
                         timesteps=200, precision=_lib.PREC_BF16X3, branches=_lib.BRANCH_COND_UNCOND, reserved=0)
    need = lib.drb_plan_workspace_bytes(C.byref(cfg))
    assert 1.0e9 < need < 4.0e9                       # 2.6 GB at the benchmark shape (z of all 15 layers is kept)
    cfg.kernel_size = 8                               # even kernels are not a 'same' convolution
    assert lib.drb_plan_workspace_bytes(C.byref(cfg)) == 0
    assert b"invalid" in lib.drb_last_error()
    cfg.kernel_size = 9; cfg.frames = 700             # more roll frames than spectrogram frames
    assert lib.drb_plan_workspace_bytes(C.byref(cfg)) == 0
    assert lib.drb_sample_loop(None, None, None, None, 0, 0, None, None) == -1


def test_python_struct_layout_matches_header():
    from diffroll_b200 import _lib
    assert C.sizeof(_lib.DrbConfig) == 16 * 4
    assert C.sizeof(_lib.DrbUpdate) == 4 + 4 + 5 * 4 + 4
    assert C.sizeof(_lib.DrbWeights) == 20 * C.sizeof(C.c_void_p)


# ---- sharding + all-gather, 2 ranks over gloo -------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gather_worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    from diffroll_b200.dist import all_gather_rolls, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(n_total * 6, dtype=torch.float32).reshape(n_total, 1, 2, 3)
    lo, hi = shard_bounds(n_total, rank, world)
    out = all_gather_rolls(full[lo:hi].clone(), n_total)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 5])
def test_all_gather_rolls_gloo_world2(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover_batch_exactly():
    from diffroll_b200.dist import shard_bounds
    for n in (1, 5, 32, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


# ---- config composition (Hydra-compatible subset) ---------------------------------------------------------------
def test_config_compose_groups_overrides_interpolation():
    from diffroll_b200.config import compose, default_config_path
    c = compose(default_config_path(), ["task=transcription", "model.args.kernel_size=9", "dataloader.batch_size=32",
                                        "task.sampling.w=0.3", "gpus=2"])
    assert c.model.name == "ClassifierFreeDiffRoll" and c.model.args.kernel_size == 9
    assert c.model.args.n_mels == 229 and c.spec.args.sample_rate == 16000 and c.spec.args.hop_length == 512
    assert c.task.sampling.type == "inpainting_ddpm_x0" and c.task.sampling.w == 0.3 and c.task.inpainting_t is None
    assert c.trainer.gpus == 2 and c.dataloader.batch_size == 32
    assert compose(default_config_path(), ["task=inpainting"]).task.inpainting_t == [500, 650]
    assert compose(default_config_path(), []).task.sampling.type == "generation_ddpm_x0"
    with pytest.raises(KeyError):
        compose(default_config_path(), ["task=does_not_exist"])
    # the composed tree constructs the model exactly like sampling.py does
    import diffroll_b200 as M
    kw = dict(c.model.args); t = dict(c.task); t.pop("name"); kw.update(t); kw["spec_args"] = dict(c.spec.args)
    m = M.ClassifierFreeDiffRoll(**kw)
    assert m.hparams.kernel_size == 9 and m.reverse_diffusion.__name__ == "inpainting_ddpm_x0"


def test_compose_reads_a_reference_style_hydra_tree(tmp_path):
    """config.compose on a directory tree laid out like the reference's config/ (defaults LIST + one file per group option,
    /root/reference/config/sampling.yaml:23-26): group files land under their group key and win over the primary file
    (Hydra 1.1 order), `group=option` and dotted overrides work, unresolvable interpolations stay lazy."""
    from diffroll_b200.config import compose
    (tmp_path / "task").mkdir()
    (tmp_path / "model").mkdir()
    (tmp_path / "sampling.yaml").write_text(
        "gpus: 1\nhop_length: 512\ntask:\n    frame_threshold: 0.8\n    extra: 7\ntrainer:\n    gpus: ${gpus}\n"
        "dataloader:\n    batch_size: 4\ndefaults:\n    - model: ClassifierFreeDiffRoll\n    - task: generation\n")
    (tmp_path / "task" / "generation.yaml").write_text(
        "name: 'generation'\nlr: ${learning_rate}\ntimesteps: 200\nframe_threshold: 0.5\nsampling:\n    type: 'generation_ddpm_x0'\n")
    (tmp_path / "task" / "transcription.yaml").write_text(
        "name: 'inpainting'\ntimesteps: 200\nframe_threshold: 0.5\nsampling:\n    type: 'inpainting_ddpm_x0'\n    w: 0.5\n")
    (tmp_path / "model" / "ClassifierFreeDiffRoll.yaml").write_text(
        "name: 'ClassifierFreeDiffRoll'\nargs:\n    kernel_size: 3\n    hop: ${hop_length}\n")
    cfg = compose(str(tmp_path / "sampling.yaml"))
    assert cfg.task.sampling.type == "generation_ddpm_x0" and cfg.task.frame_threshold == 0.5 and cfg.task.extra == 7
    assert cfg.trainer.gpus == 1 and cfg.model.args.hop == 512 and cfg.task.lr == "${learning_rate}"
    cfg = compose(str(tmp_path / "sampling.yaml"), ["task=transcription", "model.args.kernel_size=9", "dataloader.batch_size=32"])
    assert cfg.task.sampling.w == 0.5 and cfg.model.args.kernel_size == 9 and cfg.dataloader.batch_size == 32
    with pytest.raises(KeyError):
        compose(str(tmp_path / "sampling.yaml"), ["task=nope"])


def test_trainable_spec_host_side(lib_built):
    """condition='trainable_spec' (model/diffwave.py:600-605): the extra [n_mels, 641] parameter is part of the state_dict contract
    (133 tensors incl. the two mel buffers), the learned-capable plan asks for the extra clips and tables, sampling=True steps that
    cannot run as DRB_BRANCH_LEARNED run as the learned pair at guidance weight -1, and 'trainable_z' fails at construction like
    the reference (:616 vs :154)."""
    import diffroll_b200 as M
    from diffroll_b200 import _lib
    hp = default_hparams(condition="trainable_spec", sampling_type="generation_ddpm_x0")
    m = M.ClassifierFreeDiffRoll(**hp)
    assert m.trainable_parameters.shape == (229, 641) and float(m.trainable_parameters.min()) == -1.0 == float(m.trainable_parameters.max())
    sd = make_state_dict(hp)
    assert set(m.state_dict()) == set(sd) and len(sd) == 133
    m.load_state_dict(sd, strict=True)
    assert any(q is m.trainable_parameters for q in m.configure_optimizers()[0].params)
    ups, branches, _ = m._all_updates()
    assert branches == _lib.BRANCH_UNCOND
    assert m._learned_upd(ups[0], branches) is ups[0]      # shapes that run as DRB_BRANCH_LEARNED need no guidance trick
    m._learned_pair = True                                 # what _prepare decides for an odd number of 128-frame tiles
    u = m._learned_upd(ups[0], branches)
    assert u is not ups[0] and u.w == -1.0 and ups[0].w == 0.0 and list(u.s) == list(ups[0].s) and u.mode == ups[0].mode
    assert m._learned_upd(ups[0], _lib.BRANCH_COND) is ups[0]
    fixed = M.ClassifierFreeDiffRoll(**default_hparams(sampling_type="generation_ddpm_x0"))
    assert fixed._learned_upd(ups[0], _lib.BRANCH_UNCOND) is ups[0]
    spec = torch.zeros(3, 229, 128)
    out = m.uncon_dropout(spec.clone(), 0.5, mask=torch.tensor([0, 1, 0]))
    assert torch.equal(out[1], m.trainable_parameters.detach()[:, :128]) and float(out[0].abs().max()) == 0.0 == float(out[2].abs().max())
    assert float(fixed.uncon_dropout(spec.clone(), 0.5, mask=torch.tensor([0, 1, 0]))[1].max()) == -1.0
    with pytest.raises(TypeError, match="multiple values for argument 'uncond'"):
        M.ClassifierFreeDiffRoll(**default_hparams(condition="trainable_z"))
    with pytest.raises(ValueError):
        M.ClassifierFreeDiffRoll(**default_hparams(condition="something"))
    lib = _lib.load()
    cfg = _lib.DrbConfig(batch=32, frames=640, pitches=88, wave_len=327680, residual_channels=512, residual_layers=15,
                         kernel_size=9, dilation_base=2, dilation_bound=4, n_mels=229, n_fft=2048, hop_length=512,
                         timesteps=200, precision=_lib.PREC_F16N4, branches=_lib.BRANCH_COND_UNCOND, reserved=0)
    plain = lib.drb_plan_workspace_bytes(C.byref(cfg))
    cfg.branches = _lib.BRANCH_COND_LEARNED
    learned = lib.drb_plan_workspace_bytes(C.byref(cfg))
    extra = 32 * 640 * (256 * 4 + 15 * 1024 * 4)           # 32 more clips of spec32 and of every layer's conditioner table
    assert learned - plain == extra, (learned - plain, extra)
    cfg.branches = _lib.BRANCH_LEARNED                     # single-branch capacity, still with the learned clips
    single = lib.drb_plan_workspace_bytes(C.byref(cfg))
    assert 0 < single < learned
    cfg.branches = 6
    assert lib.drb_plan_workspace_bytes(C.byref(cfg)) == 0
