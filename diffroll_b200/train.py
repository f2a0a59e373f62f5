"""Host side of the training step: the reference's ``training_step`` / ``step`` in training
mode (task/diffusion.py:258-270, 651-763), the training-mode forward with spec dropout (model/diffwave.py:637-699) and
``configure_optimizers`` (torch.optim.Adam, task/diffusion.py:1057-1059), driven through the C ABI
(``drb_train_forward`` / ``drb_train_backward`` / ``drb_loss_grad`` / ``drb_adam_step``, include/diffroll_b200.h).

There is no autograd graph: the backward pass is hand-written CUDA and writes straight into ``param.grad``.  fp32
CUDA-core arithmetic, like the reference trains; CUDA tensors only, no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .diffusion_ops import LOSS_TYPES


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _param_struct(get, L):
    """DrbTrainParams from ``get(name) -> CUDA fp32 tensor`` over the reference's state_dict names.  Returns (struct, keepalive)."""
    s = _lib.DrbTrainParams()
    keep = []
    for field, name in (("in_w", "input_projection.weight"), ("in_b", "input_projection.bias"),
                        ("e1w", "diffusion_embedding.projection1.weight"), ("e1b", "diffusion_embedding.projection1.bias"),
                        ("e2w", "diffusion_embedding.projection2.weight"), ("e2b", "diffusion_embedding.projection2.bias"),
                        ("skw", "skip_projection.weight"), ("skb", "skip_projection.bias"),
                        ("hdw", "output_projection.weight"), ("hdb", "output_projection.bias")):
        setattr(s, field, get(name).data_ptr())
    for field, name in (("wd", "dilated_conv.weight"), ("bd", "dilated_conv.bias"),
                        ("wdp", "diffusion_projection.weight"), ("bdp", "diffusion_projection.bias"),
                        ("wc", "conditioner_projection.weight"), ("bc", "conditioner_projection.bias"),
                        ("wo", "output_projection.weight"), ("bo", "output_projection.bias")):
        arr = (C.c_void_p * L)(*[get(f"residual_layers.{l}.{name}").data_ptr() for l in range(L)])
        keep.append(arr)
        setattr(s, field, C.cast(arr, C.POINTER(C.c_void_p)))
    return s, keep


class TrainEngine:
    """One ``drb_train`` plan (fixed batch / frames) plus its workspace; parameters and gradients are the model's own tensors."""

    def __init__(self, model, batch, frames):
        hp = model.hparams
        self.lib = _lib.load()
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _lib.DrbError("training step: the model must live on a CUDA device; there is no CPU path")
        self.batch, self.frames = int(batch), int(frames)
        self.L = int(hp.residual_layers)
        self.cfg = _lib.DrbTrainConfig(
            batch=self.batch, frames=self.frames, pitches=88, residual_channels=int(hp.residual_channels), residual_layers=self.L,
            kernel_size=int(hp.kernel_size), dilation_base=int(hp.dilation_base), dilation_bound=int(hp.dilation_bound),
            n_mels=int(hp.spec_args["n_mels"]), timesteps=int(hp.timesteps))
        need = int(self.lib.drb_train_workspace_bytes(C.byref(self.cfg)))
        if need == 0:
            raise _lib.DrbError("drb_train_workspace_bytes: " + (self.lib.drb_last_error() or b"").decode())
        self.workspace = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        off = (-base) % 256
        plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_train_create(C.byref(plan), C.byref(self.cfg), C.c_void_p(base + off), C.c_size_t(need),
                                                 _stream(self.device)), "drb_train_create")
        self.plan = plan
        self.workspace_bytes = need
        self.spec_grad = None
        self._emb = model.diffusion_embedding.embedding.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def close(self):
        if getattr(self, "plan", None):
            self.lib.drb_train_destroy(self.plan)
            self.plan = None
            self.workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _params(self):
        named = dict(self.model.named_parameters())
        for n, q in named.items():
            if q.dtype != torch.float32 or not q.is_contiguous() or q.device != self.device:
                raise _lib.DrbError(f"training step: parameter {n} must be a contiguous fp32 tensor on {self.device}")
        return named

    def forward(self, x_t, spec, steps):
        """x_t [B,1,T,88] / [B,T,88], spec [B,n_mels,T] as the network sees it, steps int [B] -> pred [B,1,T,88]."""
        named = self._params()
        ps, keep = _param_struct(lambda n: named[n].data, self.L)
        x = x_t.to(torch.float32).reshape(self.batch, self.frames, 88).contiguous()
        sp = spec.to(torch.float32).contiguous()
        if tuple(sp.shape) != (self.batch, self.cfg.n_mels, self.frames):
            raise ValueError(f"spec must be [{self.batch}, {self.cfg.n_mels}, {self.frames}], got {tuple(sp.shape)}")
        st = steps.to(device=self.device, dtype=torch.int32).contiguous()
        pred = torch.empty(self.batch, 1, self.frames, 88, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_train_forward(self.plan, C.byref(ps), _p(x), _p(sp), _p(st), _p(self._emb), _p(pred),
                                                  _stream(self.device)), "drb_train_forward")
        self._x = x
        del keep
        return pred

    def backward(self, g_pred, accumulate=False, want_input_grad=False, want_spec_grad=False):
        """d loss / d pred [B,1,T,88] -> fills (or adds to) ``param.grad`` of every parameter; returns d loss / d x_t if asked.
        want_spec_grad: also keep d loss / d spec [B,n_mels,T] in ``self.spec_grad`` (condition='trainable_spec': the rolls
        conditioned on the learned table pass it on to that parameter)."""
        named = self._params()
        for q in named.values():
            if q.grad is None:
                q.grad = torch.zeros_like(q)
            elif not q.grad.is_contiguous():
                q.grad = q.grad.contiguous()
        ps, k1 = _param_struct(lambda n: named[n].data, self.L)
        gs, k2 = _param_struct(lambda n: named[n].grad, self.L)
        g = g_pred.to(torch.float32).reshape(self.batch, self.frames, 88).contiguous()
        gx = torch.empty_like(self._x) if want_input_grad else None
        self.spec_grad = torch.empty(self.batch, self.cfg.n_mels, self.frames, dtype=torch.float32, device=self.device) if want_spec_grad else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.drb_train_set_spec_grad(self.plan, _p(self.spec_grad) if want_spec_grad else None), "drb_train_set_spec_grad")
            _lib.check(self.lib.drb_train_backward(self.plan, C.byref(ps), C.byref(gs), _p(self._x), _p(g), C.c_int32(1 if accumulate else 0),
                                                   _p(gx) if gx is not None else None, _stream(self.device)), "drb_train_backward")
            if want_spec_grad:      # the plan must not keep a pointer into a tensor whose lifetime it does not control
                _lib.check(self.lib.drb_train_set_spec_grad(self.plan, None), "drb_train_set_spec_grad")
        del k1, k2
        return None if gx is None else gx.reshape(self.batch, 1, self.frames, 88)


def loss_grad(label, pred, loss_type, roll_scale=None):
    """d p_losses(label, pred) / d pred (task/diffusion.py:792-802), optionally times one factor per roll."""
    if loss_type not in LOSS_TYPES:
        raise NotImplementedError()
    a = label.to(torch.float32).contiguous()
    b = pred.to(torch.float32).contiguous()
    if a.shape != b.shape or not a.is_cuda:
        raise ValueError("loss_grad: label and prediction must be CUDA tensors of the same shape")
    g = torch.empty_like(b)
    rs = None if roll_scale is None else roll_scale.to(device=a.device, dtype=torch.float32).contiguous()
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.drb_loss_grad(_p(a), _p(b), _p(g), C.c_size_t(a.numel()), C.c_size_t(a.numel() // a.shape[0]),
                                     C.c_int32(LOSS_TYPES[loss_type]), _p(rs) if rs is not None else None, _stream(a.device)),
                   "drb_loss_grad")
    return g


class Adam:
    """torch.optim.Adam(params, lr) with its defaults (betas (0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad off), one fused
    kernel per tensor (``drb_adam_step``).  Same ``step()`` / ``zero_grad()`` / ``state_dict()`` surface for the part the
    reference uses (task/diffusion.py:1057-1059)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [q for q in params]
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.state = {}
        self.steps = 0

    def zero_grad(self, set_to_none=False):
        if set_to_none:
            for q in self.params:
                q.grad = None
            return
        grads = [q.grad for q in self.params if q.grad is not None]
        if grads:
            torch._foreach_zero_(grads)      # a few multi-tensor launches instead of one fill per tensor

    ADAM_CHUNK = 4096   # elements per block of drb_adam_step_multi (csrc/train.cu)

    def _tables(self, live):
        """Device tables of drb_adam_step_multi for the tensors in ``live``; rebuilt when a pointer changes."""
        key = tuple((q.data_ptr(), q.grad.data_ptr(), q.numel()) for q in live)
        if getattr(self, "_table_key", None) != key:
            rows, blocks = [], []
            for i, q in enumerate(live):
                st = self.state[id(q)]
                rows.append([q.data_ptr(), q.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), q.numel()])
                blocks += [[i, c] for c in range((q.numel() + self.ADAM_CHUNK - 1) // self.ADAM_CHUNK)]
            dev = live[0].device
            self._table = torch.tensor(rows, dtype=torch.int64).to(dev)
            self._block_map = torch.tensor(blocks, dtype=torch.int32).to(dev)
            self._table_key = key
        return self._table, self._block_map

    @torch.no_grad()
    def step(self):
        """One update of every tensor that has a gradient, as ONE launch (``drb_adam_step_multi``): torch.optim.Adam steps all of
        them together too, so they share the step count."""
        lib = _lib.load()
        self.steps += 1
        live = [q for q in self.params if q.grad is not None]
        if not live:
            return
        for q in live:
            if not q.is_cuda or q.dtype != torch.float32 or not q.is_contiguous():
                raise _lib.DrbError("Adam: parameters must be contiguous fp32 CUDA tensors")
            if not q.grad.is_contiguous():
                q.grad = q.grad.contiguous()
            st = self.state.get(id(q))
            if st is None:
                st = self.state[id(q)] = {"step": 0, "exp_avg": torch.zeros_like(q), "exp_avg_sq": torch.zeros_like(q)}
            st["step"] += 1
        steps = {self.state[id(q)]["step"] for q in live}
        dev = live[0].device
        if len(steps) == 1 and all(q.device == dev for q in live):
            table, block_map = self._tables(live)
            with torch.cuda.device(dev):
                _lib.check(lib.drb_adam_step_multi(_p(table), _p(block_map), C.c_int32(block_map.shape[0]), C.c_float(self.lr),
                                                   C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                                   C.c_float(self.weight_decay), C.c_int32(steps.pop()), _stream(dev)), "drb_adam_step_multi")
        else:       # tensors that joined later (a gradient that was None before) carry their own step count: one launch per tensor
            for q in live:
                st = self.state[id(q)]
                with torch.cuda.device(q.device):
                    _lib.check(lib.drb_adam_step(_p(q.data), _p(q.grad), _p(st["exp_avg"]), _p(st["exp_avg_sq"]), C.c_size_t(q.numel()),
                                                 C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                                 C.c_float(self.weight_decay), C.c_int32(st["step"]), _stream(q.device)), "drb_adam_step")
        # the kernels wrote through raw pointers: bump the tensors' version counters so that caches keyed on them (the
        # sampling engine's repacked weights, model._weights_version) see the update
        torch.autograd.graph.increment_version(live)
