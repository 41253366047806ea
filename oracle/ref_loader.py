"""Imports the UNMODIFIED reference (code/model.py, utils.py, loss.py, config.py) in THIS container.

Test infrastructure only.  /root/reference does not exist on the GPU box, so nothing that runs
there may call this; it is used by oracle/gen_golden.py (fixture generation) and by the
container-only differential tests (skipped when the reference checkout is absent).

The reference imports matplotlib and albumentations at module scope (utils.py:1-4,13-15,
model.py:4, config.py:1-2); neither is installed and neither is used on the hot path, so inert
stub modules are injected into sys.modules before import.
"""
import importlib
import os
import sys
import types

REF_CODE_DIRS = ["/root/reference/code"]


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()


def _stub(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Inert  # any attribute -> inert callable/class
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def available() -> bool:
    return any(os.path.isfile(os.path.join(d, "utils.py")) for d in REF_CODE_DIRS)


def load():
    """Returns (model, utils, loss, config) modules of the reference."""
    code_dir = next((d for d in REF_CODE_DIRS if os.path.isfile(os.path.join(d, "utils.py"))), None)
    if code_dir is None:
        raise RuntimeError("reference checkout not found (expected /root/reference/code)")
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
