"""ctypes binding of libdiffroll_b200.so (include/diffroll_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load,
every entry point raises.  Build it with ``python -m diffroll_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffroll_b200.so")

PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_F16F8, PREC_F16E5, PREC_F16N4 = 0, 1, 2, 3, 4, 5
BRANCH_COND_UNCOND, BRANCH_COND, BRANCH_UNCOND, BRANCH_COND_ZEROSPEC, BRANCH_COND_LEARNED, BRANCH_LEARNED = 0, 1, 2, 3, 4, 5
UPD_X0, UPD_X0_FINAL, UPD_EPS_DDPM, UPD_EPS_DDIM, UPD_EPS_FINAL, UPD_NONE = range(6)
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "f16f8": PREC_F16F8, "f16e5": PREC_F16E5,
              "f16n4": PREC_F16N4}

EXPORTS = [
    "drb_version", "drb_last_error", "drb_plan_workspace_bytes", "drb_plan_create", "drb_plan_destroy",
    "drb_plan_set_branches", "drb_plan_set_steps", "drb_time_tables", "drb_mel_forward", "drb_cond_tables", "drb_plan_use_cond_tables", "drb_in_proj", "drb_resblock_forward",
    "drb_head_posterior_step", "drb_sample_step", "drb_sample_loop", "drb_launch_count", "drb_plan_buffer",
    "drb_plan_profile", "drb_plan_profile_read", "drb_plan_profile_read2", "drb_plan_range_stats", "drb_plan_precision", "drb_extract_notes_scratch_bytes", "drb_extract_notes",
    "drb_frame_counts", "drb_q_sample", "drb_extract_x0", "drb_p_losses_scratch_bytes", "drb_p_losses", "drb_normalize_imagewise",
    "drb_train_workspace_bytes", "drb_train_create", "drb_train_destroy", "drb_train_forward", "drb_train_backward", "drb_loss_grad",
    "drb_adam_step", "drb_adam_step_multi", "drb_plan_set_step_embeddings", "drb_plan_set_uncond_spec", "drb_train_set_spec_grad",
]


class DrbConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "frames", "pitches", "wave_len", "residual_channels", "residual_layers", "kernel_size",
        "dilation_base", "dilation_bound", "n_mels", "n_fft", "hop_length", "timesteps", "precision",
        "branches", "reserved")]


_FP = C.c_void_p  # device pointers travel as integers
_FPP = C.POINTER(C.c_void_p)


class DrbWeights(C.Structure):
    _fields_ = [
        ("input_projection_w", _FP), ("input_projection_b", _FP),
        ("emb_projection1_w", _FP), ("emb_projection1_b", _FP),
        ("emb_projection2_w", _FP), ("emb_projection2_b", _FP),
        ("dilated_conv_w", _FPP), ("dilated_conv_b", _FPP),
        ("diffusion_projection_w", _FPP), ("diffusion_projection_b", _FPP),
        ("conditioner_projection_w", _FPP), ("conditioner_projection_b", _FPP),
        ("output_projection_w", _FPP), ("output_projection_b", _FPP),
        ("skip_projection_w", _FP), ("skip_projection_b", _FP),
        ("head_projection_w", _FP), ("head_projection_b", _FP),
        ("stft_window", _FP), ("mel_fb", _FP),
    ]


class DrbTrainConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "frames", "pitches", "residual_channels", "residual_layers", "kernel_size", "dilation_base", "dilation_bound",
        "n_mels", "timesteps")]


class DrbTrainParams(C.Structure):
    """Device pointers of the 132 state_dict tensors (or of their gradients), include/diffroll_b200.h drb_train_params."""
    _fields_ = [(n, _FP) for n in ("in_w", "in_b", "e1w", "e1b", "e2w", "e2b", "skw", "skb", "hdw", "hdb")] + \
               [(n, _FPP) for n in ("wd", "bd", "wdp", "bdp", "wc", "bc", "wo", "bo")]


class DrbUpdate(C.Structure):
    _fields_ = [("mode", C.c_int32), ("has_noise", C.c_int32), ("s", C.c_float * 5), ("w", C.c_float)]


class DrbError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DrbError(f"{LIB_PATH} is missing: run `python -m diffroll_b200.build` "
                       "(there is no CPU or PyTorch fallback for the sampling path)")
    lib = C.CDLL(LIB_PATH)
    lib.drb_version.restype = C.c_int
    lib.drb_last_error.restype = C.c_char_p
    lib.drb_plan_workspace_bytes.restype = C.c_size_t
    lib.drb_plan_workspace_bytes.argtypes = [C.POINTER(DrbConfig)]
    lib.drb_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(DrbConfig), C.POINTER(DrbWeights),
                                    C.c_void_p, C.c_size_t, C.c_void_p]
    lib.drb_plan_destroy.argtypes = [C.c_void_p]
    lib.drb_plan_set_branches.argtypes = [C.c_void_p, C.c_int32]
    lib.drb_time_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.drb_mel_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p]
    lib.drb_in_proj.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.drb_resblock_forward.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.drb_head_posterior_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.POINTER(DrbUpdate), C.c_void_p]
    lib.drb_sample_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(DrbUpdate), C.c_void_p]
    lib.drb_sample_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(DrbUpdate), C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]
    lib.drb_launch_count.restype = C.c_int64
    lib.drb_launch_count.argtypes = [C.c_int32]
    lib.drb_plan_buffer.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.drb_plan_profile.argtypes = [C.c_void_p, C.c_int32]
    lib.drb_plan_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.drb_plan_profile_read2.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32,
                                           C.POINTER(C.c_double), C.c_int32]
    lib.drb_plan_precision.argtypes = [C.c_void_p]
    lib.drb_plan_range_stats.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.c_void_p]
    lib.drb_extract_notes_scratch_bytes.restype = C.c_size_t
    lib.drb_extract_notes_scratch_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.drb_extract_notes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.drb_frame_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]
    lib.drb_plan_set_steps.argtypes = [C.c_void_p, C.c_void_p]
    lib.drb_cond_tables.argtypes = [C.c_void_p, C.c_void_p]
    lib.drb_plan_use_cond_tables.argtypes = [C.c_void_p, C.c_int32]
    lib.drb_plan_set_uncond_spec.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    for fn in (lib.drb_q_sample, lib.drb_extract_x0):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
    lib.drb_p_losses_scratch_bytes.restype = C.c_size_t
    lib.drb_p_losses_scratch_bytes.argtypes = []
    lib.drb_p_losses.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.drb_normalize_imagewise.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_float, C.c_void_p]
    lib.drb_plan_set_step_embeddings.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.drb_train_workspace_bytes.restype = C.c_size_t
    lib.drb_train_workspace_bytes.argtypes = [C.POINTER(DrbTrainConfig)]
    lib.drb_train_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(DrbTrainConfig), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.drb_train_destroy.argtypes = [C.c_void_p]
    lib.drb_train_destroy.restype = None
    lib.drb_train_set_spec_grad.argtypes = [C.c_void_p, C.c_void_p]
    lib.drb_train_forward.argtypes = [C.c_void_p, C.POINTER(DrbTrainParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    lib.drb_train_backward.argtypes = [C.c_void_p, C.POINTER(DrbTrainParams), C.POINTER(DrbTrainParams), C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_void_p, C.c_void_p]
    lib.drb_loss_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p]
    lib.drb_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int32, C.c_void_p]
    lib.drb_adam_step_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_int32, C.c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("drb_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().drb_last_error()
        raise DrbError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
