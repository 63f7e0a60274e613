"""SqueezeDet+ — drop-in for reference ``src/nets/squeezeDetPlus.py``: conv1 7x7/2
VALID (96) -> pool1 VALID -> fire2..4 -> pool4 -> fire5..8 -> pool8 -> fire9..11 ->
conv12 (squeeze widths 96..384: the "squeeze ratio 0.75" SqueezeNet)."""
from __future__ import annotations

from .squeezeDet import FireNetBase

# reference src/nets/squeezeDetPlus.py:40-79
_SQUEEZEDET_PLUS_BODY = (
    ('conv', 'conv1', 96, 7, 2, 'VALID'),
    ('pool', 'pool1', 3, 2, 'VALID'),
    ('fire', 'fire2', 96, 64, 64),
    ('fire', 'fire3', 96, 64, 64),
    ('fire', 'fire4', 192, 128, 128),
    ('pool', 'pool4', 3, 2, 'VALID'),
    ('fire', 'fire5', 192, 128, 128),
    ('fire', 'fire6', 288, 192, 192),
    ('fire', 'fire7', 288, 192, 192),
    ('fire', 'fire8', 384, 256, 256),
    ('pool', 'pool8', 3, 2, 'VALID'),
    ('fire', 'fire9', 384, 256, 256),
    ('fire', 'fire10', 384, 256, 256),
    ('fire', 'fire11', 384, 256, 256),
)


class SqueezeDetPlus(FireNetBase):
  BODY = _SQUEEZEDET_PLUS_BODY
