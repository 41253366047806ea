"""Constants of the reference's config module that the detection hot path reads
(code/config.py:10-11, 18-20, 37, 43-58 of GabeTsai/YOLO-For-Turbines).  The reference module also
builds Albumentations pipelines at import time; those belong to the data loader, not to this path."""
import torch

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"
BATCH_SIZE = 32

MAP_IOU_THRESHOLD = 0.5
CONF_THRESHOLD = 0.5
NMS_IOU_THRESHOLD = 0.45

DEF_IMAGE_SIZE = 416
MULTI_SCALE_TRAIN_SIZES = [416, 448, 480, 512, 544, 576, 608]

ANCHORS = [
    [(0.28, 0.22), (0.38, 0.48), (0.9, 0.78)],
    [(0.07, 0.15), (0.15, 0.11), (0.14, 0.29)],
    [(0.02, 0.03), (0.04, 0.07), (0.08, 0.06)],
]
TURBINE_ANCHORS = [
    [(0.215, 0.461), (0.992, 0.349), (0.436, 0.952)],
    [(0.06, 0.143), (0.143, 0.189), (0.408, 0.181)],
    [(0.016, 0.0349), (0.0408, 0.0598), (0.110, 0.0777)],
]
GRID_SIZES = [DEF_IMAGE_SIZE // 32, DEF_IMAGE_SIZE // 16, DEF_IMAGE_SIZE // 8]
