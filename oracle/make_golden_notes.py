"""Generate tests/golden/notes.npz with the reference's own extract_notes_wo_velocity (container only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def cases():
    rng = np.random.default_rng(7)
    T, P = 640, 88
    out = {}
    smooth = rng.random((T, P)).astype(np.float32)
    for k in range(3):                                   # make run-like structure: moving average along time
        smooth = (smooth + np.roll(smooth, 1, 0) + np.roll(smooth, 2, 0)) / 3
    out["runs"] = ((smooth - smooth.mean()) / smooth.std() * 0.25 + 0.5).astype(np.float32)
    out["noise"] = rng.random((T, P)).astype(np.float32)
    out["empty"] = np.zeros((T, P), np.float32)
    out["full"] = np.ones((T, P), np.float32)
    edge = np.zeros((T, P), np.float32); edge[0, 3] = 1; edge[T - 1, 5] = 1; edge[T - 2:, 7] = 1; edge[10:20, 87] = 1; edge[:, 0] = 1
    out["edges"] = edge
    out["short"] = rng.random((17, P)).astype(np.float32)
    out["thr"] = np.full((64, P), 0.5, np.float32); out["thr"][5:9, 1] = 0.50001
    return out


def main():
    ref_shim.import_reference()
    from task.utils import extract_notes_wo_velocity as ref_fn
    res = {}
    for name, roll in cases().items():
        p, i = ref_fn(roll, roll)
        res[f"{name}_roll"] = roll
        res[f"{name}_p"] = np.asarray(p, dtype=np.int64)
        res[f"{name}_i"] = np.asarray(i, dtype=np.int64).reshape(-1, 2)
    # separate onset / frame inputs and thresholds (the function's general form)
    rng = np.random.default_rng(11)
    on, fr = rng.random((200, 88)).astype(np.float32), rng.random((200, 88)).astype(np.float32)
    p, i = ref_fn(on, fr, onset_threshold=0.7, frame_threshold=0.4)
    res["two_on"], res["two_fr"] = on, fr
    res["two_p"], res["two_i"] = np.asarray(p, dtype=np.int64), np.asarray(i, dtype=np.int64).reshape(-1, 2)
    p, i = ref_fn(on, fr, onset_threshold=0.7, frame_threshold=0.4, rule="rule2")
    res["rule2_p"], res["rule2_i"] = np.asarray(p, dtype=np.int64), np.asarray(i, dtype=np.int64).reshape(-1, 2)
    # frame-level precision / recall / F1 as test_step computes them (task/diffusion.py:378-380), sklearn's own outputs
    import warnings
    from sklearn.metrics import precision_recall_fscore_support
    rng = np.random.default_rng(13)
    label = (rng.random((4, 1, 640, 88)) < 0.05).astype(np.float32)
    pred = np.clip(label * 0.8 + rng.normal(0, 0.35, label.shape), -1, 2).astype(np.float32)
    prf = {}
    for name, (la, pr) in {"mixed": (label, pred), "nopos": (np.zeros_like(label), pred), "nopred": (label, np.zeros_like(pred))}.items():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            p, r, f, _ = precision_recall_fscore_support(la.flatten(), pr.flatten() > 0.5, average="binary")
        prf[name] = (p, r, f)
        res[f"prf_{name}"] = np.array([p, r, f], dtype=np.float64)
    res["prf_label"], res["prf_pred"] = label, pred
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "notes.npz"), **res)
    print(prf)
    print({k: v.shape for k, v in res.items() if k.endswith("_p")})


if __name__ == "__main__":
    main()
