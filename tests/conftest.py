import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def gold():
    class G:
        nms = np.load(os.path.join(GOLD, "nms.npz"))
        nms_meta = json.load(open(os.path.join(GOLD, "nms.json")))
        iou = np.load(os.path.join(GOLD, "iou.npz"))
        decode = np.load(os.path.join(GOLD, "decode.npz"))
        map = json.load(open(os.path.join(GOLD, "map.json")))
        forward = np.load(os.path.join(GOLD, "forward.npz"))
        keys = json.load(open(os.path.join(GOLD, "state_dict_keys_nc80.json")))
        loader = json.load(open(os.path.join(GOLD, "loader.json")))
        accuracy = np.load(os.path.join(GOLD, "accuracy.npz"))
        loss = np.load(os.path.join(GOLD, "loss.npz"))
        eval_boxes = np.load(os.path.join(GOLD, "eval_boxes.npz"))
    return G


@pytest.fixture(scope="session")
def oracle_c():
    """The C restatement of the reference NMS (oracle/nms_oracle.c), built on demand."""
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    src = os.path.join(ROOT, "oracle", "nms_oracle.c")
    if not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    dll = ctypes.CDLL(so)
    dll.oracle_nms.restype = ctypes.c_int
    dll.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_int,
                               ctypes.c_void_p]

    def nms(boxes: torch.Tensor, iou_thr: float, obj_thr: float, fmt: str):
        b = boxes.detach().cpu().float().contiguous()
        keep = np.empty(max(b.shape[0], 1), dtype=np.int32)
        thr32 = float(torch.tensor(iou_thr, dtype=torch.float32))
        k = dll.oracle_nms(b.data_ptr(), b.shape[0], thr32, float(obj_thr), int(fmt == "center"), keep.ctypes.data)
        assert k >= 0
        return keep[:k].tolist()

    return nms


def template_state_dict(keys):
    return {k: torch.zeros(v, dtype=torch.long if k.endswith("num_batches_tracked") else torch.float32)
            for k, v in keys.items()}
