"""Configuration objects of the NDT1 plugin surface.

Mirrors the behaviour of the reference's ``utils/config_utils.py`` that the hot
path depends on: a dict with attribute access (``DictConfig``, config_utils.py:6-15),
recursive ``update_config(default, override)`` with ``include:<path>`` expansion
(config_utils.py:20-75) and the ``-k a.b=c`` helpers (config_utils.py:84-141).
Differences, on purpose: paths are also resolved relative to this package (the
reference resolves ``configs/ndt1.yaml`` against the CWD, models/ndt1.py:17,464),
and nested dicts are shared rather than copied on attribute access, so
``cfg.encoder.embedder.n_channels = n`` (main.py:230-231) really updates ``cfg``.
"""
from __future__ import annotations

import copy
import os
from typing import Any, Optional, Union

import yaml

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


class DictConfig(dict):
    """dict with dot access; nested dicts come back as DictConfig views."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, DictConfig):
                super().__setitem__(k, DictConfig(v))

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError as e:  # keep the reference's KeyError type (e.g. ndt1.py:481)
            raise KeyError(name) from e

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, DictConfig):
            value = DictConfig(value)
        super().__setitem__(key, value)

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def get_dict(self):
        return {k: (v.get_dict() if isinstance(v, DictConfig) else v) for k, v in self.items()}


def _resolve(path: str) -> str:
    if os.path.exists(path):
        return path
    alt = os.path.join(_PKG_DIR, path)
    if os.path.exists(alt):
        return alt
    raise FileNotFoundError(path)


def _load(path: str) -> dict:
    with open(_resolve(path), "r") as f:
        return yaml.safe_load(f)


def _unpack(cfg):
    """Expand ``include:<yaml path>`` leaves, recursively."""
    if isinstance(cfg, str) and cfg.split(":")[0] == "include":
        cfg = _load(cfg.split(":", 1)[1])
    if isinstance(cfg, dict):
        for k in list(cfg.keys()):
            cfg[k] = _unpack(cfg[k])
    return cfg


def _merge(dst, src):
    if isinstance(src, dict):
        if not isinstance(dst, dict):
            dst = {}
        for k in src:
            dst[k] = _merge(dst.get(k, {}), src[k])
        return dst
    return src


def update_config(default_config: Union[str, dict], config: Optional[Union[str, dict]] = None) -> DictConfig:
    """Values of ``config`` override ``default_config``; missing keys are added;
    either may be a yaml path; ``None`` just expands the includes."""
    if isinstance(default_config, str):
        default_config = _load(default_config)
    else:
        default_config = copy.deepcopy(default_config)
    if config is None:
        config = copy.deepcopy(default_config)
    elif isinstance(config, str):
        config = _load(config)
    else:
        config = copy.deepcopy(config)
    return DictConfig(_merge(_unpack(default_config), _unpack(config)))


def default_model_config() -> DictConfig:
    return update_config("configs/ndt1.yaml", None)


def default_trainer_config(name: str = "configs/trainer_ctc_ndt1.yaml") -> DictConfig:
    return update_config(name, None)


def convert_to_dtype(value: str):
    """String flag -> python value (config_utils.py:97-122)."""
    value = value.strip()
    if value and value[0] == "[" and value[-1] == "]":
        return [convert_to_dtype(v) for v in value[1:-1].split(",")]
    if value in ("null", "None", "none"):
        return None
    if value in ("true", "True"):
        return True
    if value in ("false", "False"):
        return False
    if value.isdigit() or value.replace("-", "").isdigit():
        return int(value)
    try:
        return float(value)
    except Exception:
        return value


def config_from_kwargs(kwargs: Optional[dict]) -> DictConfig:
    """{"a.b.c": "1"} -> {"a": {"b": {"c": 1}}} (config_utils.py:127-141)."""
    out: dict = {}
    for key, value in (kwargs or {}).items():
        value = convert_to_dtype(value) if isinstance(value, str) else value
        cur = out
        parts = key.split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = value
    return DictConfig(out)
