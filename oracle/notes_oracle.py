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
