"""`from squeezedet_b200.nets import *` mirrors reference src/nets/__init__.py:1-4."""
from .squeezeDet import SqueezeDet  # noqa: F401
from .squeezeDetPlus import SqueezeDetPlus  # noqa: F401
from .vgg16_convDet import VGG16ConvDet  # noqa: F401
from .resnet50_convDet import ResNet50ConvDet  # noqa: F401
