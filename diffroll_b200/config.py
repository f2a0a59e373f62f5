"""A small Hydra-compatible subset for sampling.py: `defaults` group selection, `${a.b}` interpolation and dotted
command-line overrides (`task=transcription model.args.kernel_size=9 task.sampling.w=0.5`).

Hydra/OmegaConf are not installed in this image; when they are, nothing stops a caller from composing the same keys
with them — the model only ever sees plain attribute dicts.
"""
from __future__ import annotations

import os
import re

import yaml

from .task import AttributeDict, to_attr

_INTERP = re.compile(r"\$\{([^}]+)\}")


def _parse_scalar(text):
    return yaml.safe_load(text)


def _get(cfg, dotted):
    cur = cfg
    for part in dotted.split("."):
        cur = cur[part]
    return cur


def _set(cfg, dotted, value):
    parts = dotted.split(".")
    cur = cfg
    for part in parts[:-1]:
        if part not in cur or not isinstance(cur[part], dict):
            cur[part] = {}
        cur = cur[part]
    cur[parts[-1]] = value


def _resolve(node, root, depth=0):
    if depth > 16:
        raise ValueError("interpolation cycle")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node)
        if m:                                   # whole-value interpolation keeps the referenced type
            return _resolve(_get(root, m.group(1)), root, depth + 1)
        return _INTERP.sub(lambda mm: str(_resolve(_get(root, mm.group(1)), root, depth + 1)), node)
    return node


def compose(config_path, overrides=()):
    """Load `config_path`, apply group selections and dotted overrides, resolve interpolations."""
    with open(config_path) as f:
        raw = yaml.safe_load(f)
    groups = raw.pop("groups", {})
    defaults = dict(raw.pop("defaults", {}))
    plain = []
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not key=value")
        key, val = ov.split("=", 1)
        key = key.lstrip("+")
        if key in groups and "." not in key:
            defaults[key] = val                 # group selection, e.g. task=transcription
        else:
            plain.append((key, _parse_scalar(val)))
    cfg = dict(raw)
    for group, choice in defaults.items():
        if group not in groups or choice not in groups[group]:
            raise KeyError(f"no option '{choice}' in config group '{group}' (have: {sorted(groups.get(group, {}))})")
        cfg[group] = yaml.safe_load(yaml.safe_dump(groups[group][choice]))  # deep copy
    for key, val in plain:
        _set(cfg, key, val)
    return to_attr(_resolve(cfg, cfg))


def default_config_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "sampling.yaml")
