"""GPU note extraction with the reference's signature (task/utils.py:4 `extract_notes_wo_velocity`).

    pitches, intervals = extract_notes_wo_velocity(roll, roll)          # roll: CUDA tensor [T, 88]
    notes = extract_notes_batch(rolls, rolls)                           # rolls [B, T, 88] -> list of (pitches, intervals)

Returns numpy arrays exactly like the reference (pitches [n], intervals [n, 2]; both empty 1-D arrays when there are no
notes).  There is no CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def extract_notes_batch(onsets, frames, onset_threshold=0.5, frame_threshold=0.5, rule="rule1"):
    if rule not in ("rule1", "rule2"):
        raise NameError("Please enter the correct rule name")
    if not (torch.is_tensor(onsets) and onsets.is_cuda and torch.is_tensor(frames) and frames.is_cuda):
        raise _lib.DrbError("extract_notes: CUDA tensors required (no CPU path)")
    if onsets.shape != frames.shape or onsets.ndim != 3:
        raise ValueError("onsets and frames must both be [B, T, P]")
    lib = _lib.load()
    on = onsets.to(torch.float32).contiguous()
    fr = on if frames is onsets else frames.to(torch.float32).contiguous()
    B, T, P = on.shape
    max_notes = T * P // 2 + P
    dev = on.device
    scratch = torch.empty(lib.drb_extract_notes_scratch_bytes(B, T, P), dtype=torch.uint8, device=dev)
    pitches = torch.empty(B, max_notes, dtype=torch.int32, device=dev)
    intervals = torch.empty(B, max_notes, 2, dtype=torch.int32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        rc = lib.drb_extract_notes(C.c_void_p(on.data_ptr()), C.c_void_p(fr.data_ptr()), B, T, P, float(onset_threshold),
                                   float(frame_threshold), 1 if rule == "rule1" else 2, C.c_void_p(scratch.data_ptr()),
                                   C.c_void_p(pitches.data_ptr()), C.c_void_p(intervals.data_ptr()),
                                   C.c_void_p(counts.data_ptr()), max_notes, stream)
    _lib.check(rc, "drb_extract_notes")
    n = counts.cpu().numpy()
    nmax = int(n.max()) if B else 0
    p_host = pitches[:, :nmax].cpu().numpy().astype(np.int64)
    i_host = intervals[:, :nmax].cpu().numpy().astype(np.int64)
    out = []
    for b in range(B):
        k = int(n[b])
        out.append((p_host[b, :k].copy(), i_host[b, :k].copy() if k else np.array([], dtype=np.float64)))
        if k == 0:
            out[-1] = (np.array([], dtype=np.float64), np.array([], dtype=np.float64))   # np.array([]) of the reference
    return out


def extract_notes_wo_velocity(onsets, frames, onset_threshold=0.5, frame_threshold=0.5, rule="rule1"):
    same = frames is onsets
    on = onsets.unsqueeze(0)
    return extract_notes_batch(on, on if same else frames.unsqueeze(0), onset_threshold, frame_threshold, rule)[0]


def frame_precision_recall_f1(label, pred, threshold=0.5):
    """precision_recall_fscore_support(label.flatten(), pred.flatten() > threshold, average='binary')[:3] of
    task/diffusion.py:378-380 from one counting kernel (positive class = 1).  Returns (precision, recall, f1, (tp, fp, fn));
    an empty denominator gives 0.0 like sklearn's zero_division default (without its warning)."""
    if not (torch.is_tensor(label) and label.is_cuda and torch.is_tensor(pred) and pred.is_cuda):
        raise _lib.DrbError("frame_precision_recall_f1: CUDA tensors required (no CPU path)")
    if label.numel() != pred.numel():
        raise ValueError("label and prediction must have the same number of elements")
    lib = _lib.load()
    la = label.to(torch.float32).contiguous()
    pr = pred.to(torch.float32).contiguous()
    counts = torch.empty(3, dtype=torch.int64, device=pr.device)
    stream = C.c_void_p(torch.cuda.current_stream(pr.device).cuda_stream)
    with torch.cuda.device(pr.device):
        _lib.check(lib.drb_frame_counts(C.c_void_p(pr.data_ptr()), C.c_void_p(la.data_ptr()), C.c_int64(pr.numel()),
                                        C.c_float(threshold), C.c_void_p(counts.data_ptr()), stream), "drb_frame_counts")
    tp, fp, fn = (int(v) for v in counts.cpu())
    precision = tp / (tp + fp) if tp + fp else 0.0
    recall = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall else 0.0
    return precision, recall, f1, (tp, fp, fn)
