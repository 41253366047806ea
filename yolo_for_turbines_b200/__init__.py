"""yolo_for_turbines_b200 -- the YOLOv3 detection hot path of GabeTsai/YOLO-For-Turbines
(forward -> anchor decode -> NMS -> mAP matching) on hand-written sm_100a CUDA kernels.

    from yolo_for_turbines_b200 import model, utils          # mirrors of the reference modules
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.utils import cells_to_boxes, non_max_suppression, calc_mAP

The native library (libyolo_b200.so, C-ABI in include/yolo_b200.h) is loaded lazily on first use
and is mandatory: nothing here falls back to ATen ops or to the CPU oracle.
"""
from . import config  # noqa: F401
from ._lib import LIB_PATH, YoloB200Error, lib  # noqa: F401

__version__ = "0.1.0"
__all__ = ["config", "model", "utils", "engine", "lib", "YoloB200Error", "LIB_PATH"]


def __getattr__(name):
    if name in ("model", "utils", "engine", "build"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
