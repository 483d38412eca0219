"""pn2_b200 -- B200 (sm_100a) implementation of the PointNet++ SA/FP geometry path and the multi-view
2D->3D feature lifting of ChengnanYu/Multi-modal-Learning-on-3D-Point-Clouds, behind the reference's
own call surface.  CUDA only: importing works anywhere, calling an operator without the built
library or without a GPU raises (no CPU fallback)."""
from . import _lib  # noqa: F401
from ._lib import Pn2Error  # noqa: F401

__all__ = ["Pn2Error"]
