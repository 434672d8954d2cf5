from .decoder import MonodepthDecoder
from .depth import DepthNet
from .encoders import create_encoder
from .pose import PoseNet

__all__ = ['DepthNet', 'PoseNet', 'MonodepthDecoder', 'create_encoder']
