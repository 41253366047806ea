"""Imports the UNMODIFIED reference (code/model.py, utils.py, loss.py, config.py).

Test infrastructure only.  Search order: baseline/_ref/code -- the offline install made by
scripts/install_reference.sh (git-ignored; travels to the GPU box with the snapshot) -- then /root/reference/code (this
container only).  Used by oracle/gen_golden*.py (fixture generation), by the container-only differential tests
(skipped when no copy is found) and by `bench.py --impl reference` / its cpu_baseline leg.  The product path never
imports this module.

The reference imports matplotlib and albumentations at module scope (utils.py:1-4,13-15,
model.py:4, config.py:1-2); neither is installed and neither is used on the hot path, so inert
stub modules are injected into sys.modules before import.
"""
import importlib
import os
import sys
import types

import torch  # noqa: F401  (before the stubs go in: torch's own import machinery inspects sys.modules)

REF_CODE_DIRS = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "code"),
                 "/root/reference/code"]


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()


def _stub(name):
    m = types.ModuleType(name)

    def _attr(attr):  # any public attribute -> inert callable/class; dunders stay missing (inspect looks at __file__)
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Inert

    m.__getattr__ = _attr
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def available() -> bool:
    return any(os.path.isfile(os.path.join(d, "utils.py")) for d in REF_CODE_DIRS)


def load():
    """Returns (model, utils, loss, config) modules of the reference."""
    code_dir = next((d for d in REF_CODE_DIRS if os.path.isfile(os.path.join(d, "utils.py"))), None)
    if code_dir is None:
        raise RuntimeError("reference not found (looked in baseline/_ref/code and /root/reference/code)")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "albumentations",
                 "albumentations.pytorch"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    if code_dir not in sys.path:
        sys.path.insert(0, code_dir)
    saved = {k: sys.modules.pop(k) for k in ("config", "utils", "model", "loss") if k in sys.modules}
    try:
        cfg = importlib.import_module("config")
        utils = importlib.import_module("utils")
        model = importlib.import_module("model")
        loss = importlib.import_module("loss")
    finally:
        # keep the reference modules reachable only through the returned handles
        for k in ("config", "utils", "model", "loss"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
        if code_dir in sys.path:
            sys.path.remove(code_dir)
    return model, utils, loss, cfg
