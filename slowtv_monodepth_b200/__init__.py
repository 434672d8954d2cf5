"""slowtv_monodepth_b200 — B200-native (sm_100a) training hot path of jspenmar/slowtv_monodepth.

Host side mirrors the reference's plugin surface (src/registry.py, src.losses, src.regularizers, src.networks,
src.tools.geometry, src.core.handlers); device side is hand-written CUDA behind the C ABI in `include/stv.h`.
"""
__version__ = '0.1.0'
