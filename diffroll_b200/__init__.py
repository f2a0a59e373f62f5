"""diffroll_b200 — B200-native sampling hot path of sony/DiffRoll.

``import diffroll_b200 as Model; getattr(Model, cfg.model.name)`` resolves
``ClassifierFreeDiffRoll`` the way the reference's ``sampling.py:54`` does with its own
``model`` package.  Importing this package does not need a GPU or the shared library;
running anything does (there is no CPU fallback).
"""
from .model import ClassifierFreeDiffRoll, Normalization  # noqa: F401
from .task import SpecRollDiffusion  # noqa: F401

__all__ = ["ClassifierFreeDiffRoll", "SpecRollDiffusion", "Normalization"]
__version__ = "0.1.0"
