"""Load the reference's own, unmodified ``model.py`` / ``data.py`` / ``training.py`` / ``constants.py``.

TEST INFRASTRUCTURE ONLY. Works only where ``/root/reference`` exists (the build container, never the
GPU box). Nothing is copied: the files are imported in place through ``sys.path`` on top of
``oracle.pyg_shim``. Used by ``tests/golden/make_golden.py`` and by the container-only tests that pin
``oracle.graph_oracle`` / ``oracle.model_oracle`` against the real reference.
"""
from __future__ import annotations

import importlib
import os
import sys
from types import SimpleNamespace

REFERENCE_DIR = os.environ.get("POLYPHEMUS_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


_cache = None


def load():
    """Returns a namespace with the reference modules: constants, data, model, training."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    from . import pyg_shim

    pyg_shim.install()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    cwd = os.getcwd()
    try:
        os.chdir(REFERENCE_DIR)  # generation_config.py:5,15 opens a relative yaml at import of utils
        mods = {}
        for name in ("constants", "data", "model"):
            mods[name] = importlib.import_module(name)
        try:
            mods["training"] = importlib.import_module("training")
        except Exception as exc:  # tqdm etc. — training is only needed for _losses
            mods["training"] = None
            mods["training_error"] = exc
    finally:
        os.chdir(cwd)
    _cache = SimpleNamespace(**mods)
    return _cache
