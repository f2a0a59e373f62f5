"""CPU oracle for the post-loop decode: piano roll -> note list.

TEST INFRASTRUCTURE ONLY (same rules as oracle/diffroll_oracle.py).

Restates ``extract_notes_wo_velocity`` of the reference (task/utils.py:4-54, duplicated at
task/diffusion.py:1185-1235) with numpy; the reference calls it on every finished roll with
``onsets = frames = roll`` (task/diffusion.py:599-602).  Parity status: PINNED — tests/golden/notes.npz holds
outputs of the reference's own function (imported through oracle/ref_shim.py by oracle/make_golden_notes.py).
Integer work: the bar is bit-exact, including the order of the notes (np.nonzero order: frame-major, then pitch).
"""
from __future__ import annotations

import numpy as np


def extract_notes_wo_velocity(onsets, frames, onset_threshold=0.5, frame_threshold=0.5, rule="rule1"):
    onsets = (np.asarray(onsets) > onset_threshold).astype(int)
    frames = (np.asarray(frames) > frame_threshold).astype(int)
    onset_diff = np.concatenate([onsets[:1, :], onsets[1:, :] - onsets[:-1, :]], axis=0) == 1
    if rule == "rule2":
        pass
    elif rule == "rule1":
        onset_diff = onset_diff & (frames == 1)
    else:
        raise NameError("Please enter the correct rule name")
    pitches, intervals = [], []
    n_frames = onsets.shape[0]
    for frame, pitch in zip(*np.nonzero(onset_diff)):
        offset = frame
        while onsets[offset, pitch] or frames[offset, pitch]:
            offset += 1
            if offset == n_frames:
                break
        if offset > frame:
            pitches.append(pitch)
            intervals.append([frame, offset])
    return np.array(pitches), np.array(intervals)


def frame_precision_recall_f1(label, pred, threshold=0.5):
    """precision_recall_fscore_support(label.flatten(), pred.flatten() > threshold, average='binary') of
    task/diffusion.py:378-380 (third-party: scikit-learn, requirements.txt) restated from its published definition for
    the binary average with pos_label = 1: P = TP/(TP+FP), R = TP/(TP+FN), F1 = 2PR/(P+R), 0 on an empty denominator.
    Pinned by tests/golden/notes.npz (sklearn's own outputs, oracle/make_golden_notes.py)."""
    y = np.asarray(label).reshape(-1) == 1
    p = np.asarray(pred).reshape(-1) > threshold
    tp, fp, fn = int((p & y).sum()), int((p & ~y).sum()), int((~p & y).sum())
    prec = tp / (tp + fp) if tp + fp else 0.0
    rec = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
    return prec, rec, f1, (tp, fp, fn)
