"""`from squeezedet_b200.config import *` mirrors reference src/config/__init__.py:1-5."""
from .config import (  # noqa: F401
    ModelConfig, base_model_config, make_anchor_box, set_anchors,
    kitti_squeezeDet_config, kitti_squeezeDetPlus_config, kitti_vgg16_config,
    kitti_res50_config,
)
