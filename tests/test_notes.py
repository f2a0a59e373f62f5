"""Post-loop decode (row f2): oracle vs the reference's own outputs (CPU), CUDA kernels vs oracle (GPU), bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle.notes_oracle import extract_notes_wo_velocity as oracle_notes

CASES = ["runs", "noise", "empty", "full", "edges", "short", "thr"]


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.size == b.size and (a.size == 0 or np.array_equal(a.reshape(b.shape), b))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = golden("notes.npz")
    p, i = oracle_notes(g[f"{name}_roll"], g[f"{name}_roll"])
    assert _same(p, g[f"{name}_p"]) and _same(i, g[f"{name}_i"])


def test_oracle_two_inputs_and_thresholds():
    g = golden("notes.npz")
    p, i = oracle_notes(g["two_on"], g["two_fr"], onset_threshold=0.7, frame_threshold=0.4)
    assert _same(p, g["two_p"]) and _same(i, g["two_i"])
    p, i = oracle_notes(g["two_on"], g["two_fr"], onset_threshold=0.7, frame_threshold=0.4, rule="rule2")
    assert _same(p, g["rule2_p"]) and _same(i, g["rule2_i"]) and len(g["rule2_p"]) > len(g["two_p"])
    with pytest.raises(NameError):
        oracle_notes(g["two_on"], g["two_fr"], rule="rule3")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_notes_vs_golden(name):
    from diffroll_b200.notes import extract_notes_wo_velocity
    g = golden("notes.npz")
    roll = torch.from_numpy(g[f"{name}_roll"]).cuda()
    p, i = extract_notes_wo_velocity(roll, roll)
    assert _same(p, g[f"{name}_p"]) and _same(i, g[f"{name}_i"])


@pytest.mark.gpu
def test_cuda_notes_two_inputs_batch_and_full_size():
    from diffroll_b200.notes import extract_notes_batch, extract_notes_wo_velocity
    g = golden("notes.npz")
    p, i = extract_notes_wo_velocity(torch.from_numpy(g["two_on"]).cuda(), torch.from_numpy(g["two_fr"]).cuda(), 0.7, 0.4)
    assert _same(p, g["two_p"]) and _same(i, g["two_i"])
    p, i = extract_notes_wo_velocity(torch.from_numpy(g["two_on"]).cuda(), torch.from_numpy(g["two_fr"]).cuda(), 0.7, 0.4, rule="rule2")
    assert _same(p, g["rule2_p"]) and _same(i, g["rule2_i"])
    with pytest.raises(NameError):
        extract_notes_wo_velocity(torch.from_numpy(g["two_on"]).cuda(), torch.from_numpy(g["two_fr"]).cuda(), rule="rule3")
    # BASELINE configs[1] size: 32 rolls of 640 x 88, against the oracle roll by roll
    rng = np.random.default_rng(3)
    rolls = rng.random((32, 640, 88)).astype(np.float32)
    rolls[5] = 0.0; rolls[6] = 1.0
    out = extract_notes_batch(torch.from_numpy(rolls).cuda(), torch.from_numpy(rolls).cuda(), 0.6, 0.6)
    for b in (0, 5, 6, 31):
        p, i = oracle_notes(rolls[b], rolls[b], 0.6, 0.6)
        assert _same(out[b][0], p) and _same(out[b][1], i)
    # properties at full size: sorted by (onset, pitch); offsets after onsets; every note starts on a rising edge
    for b in range(32):
        p, i = out[b]
        if len(p) == 0:
            continue
        key = i[:, 0] * 128 + p
        assert np.all(np.diff(key) > 0) and np.all(i[:, 1] > i[:, 0]) and i[:, 1].max() <= 640
        on = rolls[b] > 0.6
        assert np.all(on[i[:, 0], p]) and np.all((i[:, 0] == 0) | ~on[np.maximum(i[:, 0] - 1, 0), p])


PRF_CASES = ["mixed", "nopos", "nopred"]


def _prf_inputs(g, name):
    label, pred = g["prf_label"], g["prf_pred"]
    if name == "nopos":
        label = np.zeros_like(label)
    if name == "nopred":
        pred = np.zeros_like(pred)
    return label, pred


@pytest.mark.parametrize("name", PRF_CASES)
def test_oracle_frame_prf_matches_sklearn_golden(name):
    """test_step's frame metrics (task/diffusion.py:378-380): the restated definition against sklearn's own outputs."""
    from oracle.notes_oracle import frame_precision_recall_f1
    g = golden("notes.npz")
    label, pred = _prf_inputs(g, name)
    p, r, f, _ = frame_precision_recall_f1(label, pred, 0.5)
    assert np.allclose([p, r, f], g[f"prf_{name}"], rtol=0, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PRF_CASES)
def test_cuda_frame_prf_vs_golden_and_oracle(name):
    from diffroll_b200.notes import frame_precision_recall_f1 as cuda_prf
    from oracle.notes_oracle import frame_precision_recall_f1 as oracle_prf
    g = golden("notes.npz")
    label, pred = _prf_inputs(g, name)
    p, r, f, counts = cuda_prf(torch.from_numpy(label).cuda(), torch.from_numpy(pred).cuda(), 0.5)
    assert counts == oracle_prf(label, pred, 0.5)[3]                       # integer counts: bit-exact
    assert np.allclose([p, r, f], g[f"prf_{name}"], rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_test_step_end_to_end_short_chain():
    """test_step (task/diffusion.py:312-418 up to the mir_eval boundary): sampling from fresh noise, frame metrics and
    note lists; the metrics must equal the oracle's on the returned roll, the label notes the oracle's on the label."""
    import diffroll_b200 as M
    from diffroll_b200.synthetic import default_hparams, make_labelled_batch, make_state_dict
    from oracle.notes_oracle import frame_precision_recall_f1 as oracle_prf
    hp = default_hparams(timesteps=6)
    m = M.ClassifierFreeDiffRoll(**hp)
    m.load_state_dict(make_state_dict(hp))
    m = m.cuda().eval()
    frame, audio, _, _ = make_labelled_batch(B=2)
    torch.manual_seed(3)
    out = m.test_step({"frame": frame.cuda(), "audio": audio.cuda()}, 0)
    assert out["roll_pred"].shape == (2, 1, 128, 88)
    p, r, f, counts = oracle_prf(frame.numpy(), out["roll_pred"][:, 0], 0.5)
    assert counts == out["counts"] and (p, r, f) == (out["frame_p"], out["frame_r"], out["frame_f1"])
    for b in range(2):
        pr, ir = oracle_notes(frame[b].numpy(), frame[b].numpy())
        assert _same(out["notes_ref"][b][0], pr) and _same(out["notes_ref"][b][1], ir)
        pe, ie = oracle_notes(out["roll_pred"][b, 0], out["roll_pred"][b, 0])
        assert _same(out["notes_est"][b][0], pe) and _same(out["notes_est"][b][1], ie)
    for e, _ in m._engines.values():
        e.close()
