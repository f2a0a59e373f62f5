"""Import the UNMODIFIED reference (`/root/reference`) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path.

The reference needs pytorch_lightning / hydra / matplotlib / mir_eval / mido,
none of which are installed here.  None of them does arithmetic
on the sampling path, so this module installs inert stand-ins into
``sys.modules`` and then imports ``model`` from the reference tree.  The only
stand-in with behaviour is ``pl.LightningModule``: an ``nn.Module`` whose
``save_hyperparameters()`` records the constructor arguments of every
``__init__`` frame of ``self`` (what Lightning 1.6 does), because the reference
reads ``self.hparams.timesteps``, ``.condition``, ``.sampling.w`` … later
(task/diffusion.py:235,528,1007-1009; model/diffwave.py:647,657).

``/root/reference`` does not exist on the GPU box; this file is used only by
``oracle/make_golden.py`` and by the container-only test that pins
``oracle/diffroll_oracle.py`` against the real reference.
"""
from __future__ import annotations

import importlib
import inspect
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DIFFROLL_REFERENCE", "/root/reference")


class AttrDict(dict):
    """dict with attribute access, enough of OmegaConf's DictConfig for the reference."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        hp = AttrDict()
        frame = inspect.currentframe().f_back
        # walk outwards over every __init__ frame whose `self` is this object
        frames = []
        while frame is not None:
            if frame.f_code.co_name == "__init__" and frame.f_locals.get("self") is self:
                frames.append(frame)
            frame = frame.f_back
        for fr in reversed(frames):  # outermost (subclass) first, base class last wins
            info = inspect.getargvalues(fr)
            for name in info.args:
                if name != "self":
                    hp[name] = fr.f_locals[name]
            if info.keywords:
                hp.update(fr.f_locals[info.keywords])
        hp.pop("kwargs", None)
        object.__setattr__(self, "_hparams", hp)

    @property
    def hparams(self):
        return self._hparams

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        return _Anything()


def install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = _module("pytorch_lightning", LightningModule=_LightningModule, Trainer=_Anything)
        _module("pytorch_lightning.callbacks", LearningRateMonitor=_Anything, ModelCheckpoint=_Anything)
        _module("pytorch_lightning.loggers", TensorBoardLogger=_Anything)
        pl.callbacks = sys.modules["pytorch_lightning.callbacks"]
        pl.loggers = sys.modules["pytorch_lightning.loggers"]
    for name, attrs in [
        ("matplotlib", {}),
        ("matplotlib.pyplot", {}),
        ("matplotlib.animation", {}),
        ("mpl_toolkits", {}),
        ("mpl_toolkits.axes_grid1", {"make_axes_locatable": _Anything()}),
        ("mir_eval", {}),
        ("mir_eval.transcription", {"precision_recall_f1_overlap": _Anything()}),
        ("mir_eval.util", {"midi_to_hz": _Anything(), "hz_to_midi": _Anything()}),
        ("mido", {"Message": _Anything, "MidiFile": _Anything, "MidiTrack": _Anything}),
    ]:
        try:
            importlib.import_module(name)
        except Exception:
            _module(name, **attrs)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def import_reference():
    """Return the reference's ``model`` package (unmodified source)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for clash in ("model", "task", "utils"):
        mod = sys.modules.get(clash)
        if mod is not None and not (getattr(mod, "__file__", None) or "").startswith(REFERENCE_ROOT):
            del sys.modules[clash]
    return importlib.import_module("model")


def build_reference_model(hp: dict):
    """Construct the reference ``ClassifierFreeDiffRoll`` from a plain hparam dict."""
    Model = import_reference()
    kw = dict(hp)
    for k in ("spec_args", "training", "sampling"):
        kw[k] = to_attr(kw[k])
    m = Model.ClassifierFreeDiffRoll(**kw)
    m.eval()
    return m
