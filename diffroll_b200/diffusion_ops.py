"""ctypes wrappers of the elementwise / reduction kernels around the network forward in the reference's
(validation) step, task/diffusion.py:651-763: q_sample (:31-46), extract_x0 (:49-65), p_losses (:792-802) and the
label-roll Normalization (model/utils.py:21-32).  CUDA tensors only: there is no CPU path."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

LOSS_TYPES = {"l1": 0, "l2": 1, "huber": 2}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise _lib.DrbError(f"{name}: diffroll_b200 runs on CUDA tensors only; there is no CPU path")
    return t.to(torch.float32).contiguous()


def _diffuse(fn, a, b, t, sa, s1, what):
    a, b = _f32(a, what), _f32(b, what)
    if a.shape != b.shape or a.ndim < 2:
        raise ValueError(f"{what}: operands must have the same [B, ...] shape")
    B = a.shape[0]
    n_per = a.numel() // B
    steps = t.to(device=a.device, dtype=torch.int32).contiguous()
    if tuple(steps.shape) != (B,):
        raise ValueError(f"{what}: t must hold one step per roll")
    sa_d = sa.to(device=a.device, dtype=torch.float32).contiguous()
    s1_d = s1.to(device=a.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(a)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(getattr(lib, fn)(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(steps.data_ptr()),
                                    C.c_void_p(sa_d.data_ptr()), C.c_void_p(s1_d.data_ptr()), C.c_void_p(out.data_ptr()),
                                    C.c_int32(B), C.c_int64(n_per), _stream()), fn)
    return out


def q_sample(x_start, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, noise=None):
    """task/diffusion.py:31-46 (same argument order): sqrt(abar_t) * x_start + sqrt(1 - abar_t) * noise, t per roll."""
    if noise is None:
        noise = torch.randn_like(x_start)
    return _diffuse("drb_q_sample", x_start, noise, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, "q_sample")


def extract_x0(x_t, epsilon, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod):
    """task/diffusion.py:49-65: (x_t - sqrt(1 - abar_t) * epsilon) / sqrt(abar_t)."""
    return _diffuse("drb_extract_x0", x_t, epsilon, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, "extract_x0")


def p_losses(label, prediction, loss_type="l1"):
    """task/diffusion.py:792-802: F.l1_loss / F.mse_loss / F.smooth_l1_loss (mean) as a 0-dim CUDA tensor."""
    if loss_type not in LOSS_TYPES:
        raise NotImplementedError()                     # task/diffusion.py:800
    a, b = _f32(label, "p_losses"), _f32(prediction, "p_losses")
    if a.shape != b.shape:
        raise ValueError("p_losses: label and prediction must have the same shape")
    lib = _lib.load()
    lib.drb_p_losses_scratch_bytes.restype = C.c_size_t
    scratch = torch.empty(int(lib.drb_p_losses_scratch_bytes()), dtype=torch.uint8, device=a.device)
    out = torch.empty((), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.drb_p_losses(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_int64(a.numel()),
                                    C.c_int32(LOSS_TYPES[loss_type]), C.c_void_p(scratch.data_ptr()),
                                    C.c_void_p(out.data_ptr()), _stream()), "drb_p_losses")
    return out


def normalize_imagewise(x, lo=0.0, hi=1.0):
    """model/utils.py:21-32 ('imagewise'): per-roll min-max to [lo, hi]; an empty (constant) roll becomes lo."""
    a = _f32(x, "normalize")
    B = a.shape[0]
    out = torch.empty_like(a)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.drb_normalize_imagewise(C.c_void_p(a.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int32(B),
                                               C.c_int64(a.numel() // B), C.c_float(lo), C.c_float(hi), _stream()),
                   "drb_normalize_imagewise")
    return out
