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


def _resolve(node, root, depth=0, lazy=False):
    """Resolve `${a.b}` interpolations.  lazy=True leaves an interpolation whose target does not exist as the literal string,
    like OmegaConf, which only fails when such a key is READ (the reference's task files carry `lr: ${learning_rate}`, a
    training-only key that config/sampling.yaml never defines and sampling.py never reads)."""
    if depth > 16:
        raise ValueError("interpolation cycle")
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth, lazy) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth, lazy) for v in node]
    if isinstance(node, str):
        def look(path):
            return _resolve(_get(root, path), root, depth + 1, lazy)
        m = _INTERP.fullmatch(node)
        try:
            if m:                                   # whole-value interpolation keeps the referenced type
                return look(m.group(1))
            return _INTERP.sub(lambda mm: str(look(mm.group(1))), node)
        except (KeyError, TypeError):
            if lazy:
                return node
            raise
    return node


def _deep_merge(base, top):
    """`top` over `base`, dictionaries merged key by key (lists and scalars replaced)."""
    out = dict(base)
    for k, v in top.items():
        out[k] = _deep_merge(out[k], v) if isinstance(v, dict) and isinstance(out.get(k), dict) else v
    return out


def _compose_tree(config_path, raw, overrides):
    """Reference-style Hydra tree (config/sampling.yaml:23-26 there): `defaults` is a LIST of {group: option} and every
    option is a file `<dir>/<group>/<option>.yaml` whose content lands under the key `<group>`.  The reference pins
    hydra-core 1.2.0 and calls `@hydra.main` without `version_base`, i.e. with Hydra 1.1's composition order: the primary
    file comes FIRST and the group files are merged over it (task/generation.yaml's `frame_threshold: 0.5` wins over the
    primary's 0.8)."""
    root_dir = os.path.dirname(os.path.abspath(config_path))
    defaults = raw.pop("defaults")
    choice, order = {}, []
    for ent in defaults:
        if ent == "_self_":
            continue
        if not isinstance(ent, dict) or len(ent) != 1:
            raise ValueError(f"unsupported defaults entry {ent!r}")
        (g, opt), = ent.items()
        g = str(g).replace("override ", "").strip()
        choice[g] = opt
        order.append(g)
    plain = []
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not key=value")
        key, val = ov.split("=", 1)
        added = key.startswith("+")
        key = key.lstrip("+")
        if "." not in key and (key in choice or (added and os.path.isdir(os.path.join(root_dir, key)))):
            if key not in choice:
                order.append(key)
            choice[key] = val                   # group selection, e.g. task=transcription
        else:
            plain.append((key, _parse_scalar(val)))
    cfg = dict(raw)
    for g in order:
        path = os.path.join(root_dir, g, f"{choice[g]}.yaml")
        if not os.path.exists(path):
            have = sorted(f[:-5] for f in os.listdir(os.path.join(root_dir, g)) if f.endswith(".yaml")) if os.path.isdir(os.path.join(root_dir, g)) else []
            raise KeyError(f"no option '{choice[g]}' in config group '{g}' (have: {have})")
        with open(path) as f:
            body = yaml.safe_load(f) or {}
        cfg = _deep_merge(cfg, {g: body})
    for key, val in plain:
        _set(cfg, key, val)
    return to_attr(_resolve(cfg, cfg, lazy=True))


def compose(config_path, overrides=()):
    """Load `config_path`, apply group selections and dotted overrides, resolve interpolations.  Two layouts are read:
    this repository's single file with a `groups:` section, and the reference's Hydra directory tree (`defaults:` list +
    one file per group option)."""
    with open(config_path) as f:
        raw = yaml.safe_load(f)
    if isinstance(raw.get("defaults"), list):   # a reference checkout's config directory drops in unchanged
        return _compose_tree(config_path, raw, overrides)
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
