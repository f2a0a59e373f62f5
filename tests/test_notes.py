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
