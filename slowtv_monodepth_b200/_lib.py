"""ctypes binding of libstv.so (the C ABI declared in include/stv.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the product path raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

# STV_LIB: developer override used to A/B-test differently tuned builds of the same sources.
LIB_PATH = Path(os.environ.get('STV_LIB') or Path(__file__).resolve().parent/'libstv.so')
MAX_SCALES = 8
SEL_STATIC, SEL_MEAN = 255, 254

_lib = None


class StvError(RuntimeError):
    pass


class PhotoCfg(C.Structure):
    _fields_ = [('b', C.c_int), ('n', C.c_int), ('S', C.c_int), ('H', C.c_int), ('W', C.c_int),
                ('w_ssim', C.c_float), ('w_l1', C.c_float), ('use_min', C.c_int), ('use_automask', C.c_int),
                ('noise_seed', C.c_uint64), ('depth_stride_s', C.c_int64)]


class ReconCfg(C.Structure):
    _fields_ = [('b', C.c_int), ('n', C.c_int), ('C', C.c_int), ('H', C.c_int), ('W', C.c_int), ('loss', C.c_int),
                ('use_min', C.c_int), ('use_automask', C.c_int), ('mask_mode', C.c_int), ('noise_seed', C.c_uint64)]


RECON_LOSS = {'ssim': 0, 'l1': 1, 'l2': 2}
RECON_MASK = {None: 0, 'explainability': 1, 'uncertainty': 2}
REGR_LOSS = {'l1': 0, 'log_l1': 1, 'berhu': 2}


class PhotoSrc(C.Structure):
    _fields_ = [('mode', C.c_int), ('h', C.c_int*MAX_SCALES), ('w', C.c_int*MAX_SCALES), ('min_depth', C.c_float), ('max_depth', C.c_float)]


class GemmEpi(C.Structure):
    _fields_ = [('bias', C.c_void_p), ('aux', C.c_void_p), ('gamma', C.c_void_p), ('res', C.c_void_p), ('dact_src', C.c_void_p),
                ('colsum', C.c_void_p), ('act', C.c_int), ('dact', C.c_int), ('accumulate', C.c_int)]


ACT = {None: 0, 'none': 0, 'relu': 1, 'gelu': 2, 'elu': 3, 'sigmoid': 4}


class ConvGeom(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('N', 'H', 'W', 'C1', 'C2', 'up1', 'Cout', 'R', 'S', 'stride', 'pad', 'reflect')]


class SmoothCfg(C.Structure):
    _fields_ = [('b', C.c_int), ('S', C.c_int), ('H', C.c_int), ('W', C.c_int),
                ('h', C.c_int*MAX_SCALES), ('w', C.c_int*MAX_SCALES), ('scale_div', C.c_float*MAX_SCALES),
                ('use_edges', C.c_int)]


_P = C.c_void_p
_SIGNATURES = {
    'stv_version': (C.c_int, []),
    'stv_last_error': (C.c_char_p, []),
    'stv_launch_count': (C.c_ulonglong, []),
    'stv_photo_workspace_bytes': (C.c_size_t, [C.POINTER(PhotoCfg)]),
    'stv_photo_fwd': (C.c_int, [C.POINTER(PhotoCfg)] + [_P]*12 + [C.c_size_t, _P]),
    'stv_photo_bwd': (C.c_int, [C.POINTER(PhotoCfg)] + [_P]*13 + [C.c_size_t, _P]),
    'stv_photo_error': (C.c_int, [C.POINTER(PhotoCfg), _P, _P, _P, _P]),
    'stv_photo_fused_workspace_bytes': (C.c_size_t, [C.POINTER(PhotoCfg)]),
    'stv_photo_fused_partial_bytes': (C.c_size_t, [C.POINTER(PhotoCfg)]),
    'stv_photo_fused_fwd': (C.c_int, [C.POINTER(PhotoCfg), C.POINTER(PhotoSrc), _P, _P, _P, C.c_ulonglong] + [_P]*11 + [C.c_size_t, _P]),
    'stv_photo_fused_bwd_workspace_bytes': (C.c_size_t, [C.POINTER(PhotoCfg), C.POINTER(PhotoSrc)]),
    'stv_photo_fused_bwd': (C.c_int, [C.POINTER(PhotoCfg), C.POINTER(PhotoSrc)] + [_P]*10 + [C.c_size_t, _P]),
    'stv_tex_create': (C.c_int, [_P, C.c_longlong, C.c_int, C.POINTER(C.c_ulonglong)]),
    'stv_tex_destroy': (C.c_int, [C.c_ulonglong]),
    'stv_recon_workspace_bytes': (C.c_size_t, [C.POINTER(PhotoCfg)]),
    'stv_recon_fwd': (C.c_int, [C.POINTER(PhotoCfg)] + [_P]*9 + [C.c_size_t, _P]),
    'stv_recon_bwd': (C.c_int, [C.POINTER(PhotoCfg)] + [_P]*6),
    'stv_regr_workspace_bytes': (C.c_size_t, []),
    'stv_regr_fwd': (C.c_int, [C.c_longlong, C.c_int, C.c_int] + [_P]*6 + [C.c_size_t, _P]),
    'stv_regr_bwd': (C.c_int, [C.c_longlong, C.c_int, C.c_int] + [_P]*7 + [C.c_size_t, _P]),
    'stv_pwreg_fwd': (C.c_int, [C.c_longlong, C.c_int, C.c_float] + [_P]*3 + [C.c_size_t, _P]),
    'stv_pwreg_bwd': (C.c_int, [C.c_longlong, C.c_int, C.c_float] + [_P]*4),
    'stv_feat_reg_workspace_bytes': (C.c_size_t, [C.c_int]*5),
    'stv_feat_reg_fwd': (C.c_int, [C.c_int]*7 + [_P]*5 + [C.c_size_t, _P]),
    'stv_feat_reg_bwd': (C.c_int, [C.c_int]*7 + [_P]*4 + [C.c_size_t, _P]),
    'stv_smooth_ex_workspace_bytes': (C.c_size_t, [C.c_int]*4),
    'stv_smooth_ex_fwd': (C.c_int, [C.c_int]*7 + [_P]*6 + [C.c_size_t, _P]),
    'stv_smooth_ex_bwd': (C.c_int, [C.c_int]*7 + [_P]*4 + [C.c_size_t, _P]),
    'stv_recon_ex_workspace_bytes': (C.c_size_t, [C.POINTER(ReconCfg)]),
    'stv_recon_ex_fwd': (C.c_int, [C.POINTER(ReconCfg)] + [_P]*10 + [C.c_size_t, _P]),
    'stv_recon_ex_bwd': (C.c_int, [C.POINTER(ReconCfg)] + [_P]*9),
    'stv_view_synth_fwd': (C.c_int, [C.c_int]*4 + [_P]*9),
    'stv_view_synth_workspace_bytes': (C.c_size_t, [C.c_int]*4),
    'stv_view_synth_bwd': (C.c_int, [C.c_int]*4 + [_P]*13 + [C.c_size_t, _P]),
    'stv_disp_to_depth_fwd': (C.c_int, [C.c_int]*5 + [C.c_float]*2 + [_P]*4),
    'stv_disp_to_depth_bwd_workspace_bytes': (C.c_size_t, [C.c_int]*5),
    'stv_disp_to_depth_bwd': (C.c_int, [C.c_int]*5 + [C.c_float]*2 + [_P]*5 + [C.c_size_t, _P]),
    'stv_smooth_workspace_bytes': (C.c_size_t, [C.POINTER(SmoothCfg)]),
    'stv_smooth_fwd': (C.c_int, [C.POINTER(SmoothCfg), _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'stv_smooth_bwd': (C.c_int, [C.POINTER(SmoothCfg), _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'stv_dwconv7_fwd': (C.c_int, [C.c_int]*4 + [_P]*5 + [C.c_int, _P]),
    'stv_dwconv7_wgrad_workspace_bytes': (C.c_size_t, [C.c_int]*4),
    'stv_dwconv7_wgrad': (C.c_int, [C.c_int]*4 + [_P]*4 + [C.c_int, _P, C.c_size_t, _P]),
    'stv_layernorm_fwd': (C.c_int, [C.c_longlong, C.c_int, _P, _P, _P, C.c_float, _P, _P, _P, _P]),
    'stv_layernorm_bwd_workspace_bytes': (C.c_size_t, [C.c_longlong, C.c_int]),
    'stv_layernorm_bwd': (C.c_int, [C.c_longlong, C.c_int] + [_P]*8 + [C.c_int, _P, C.c_size_t, _P]),
    'stv_gemm_tf32': (C.c_int, [C.c_int]*3 + [_P, C.c_longlong, C.c_int, _P, C.c_longlong, C.c_int, _P, C.c_longlong,
                                C.POINTER(GemmEpi), C.c_int, _P]),
    'stv_conv_fprop': (C.c_int, [C.POINTER(ConvGeom), _P, _P, _P, _P, C.POINTER(GemmEpi), _P]),
    'stv_conv_dgrad': (C.c_int, [C.POINTER(ConvGeom), _P, _P, _P, C.POINTER(GemmEpi), _P]),
    'stv_conv_wgrad': (C.c_int, [C.POINTER(ConvGeom), _P, _P, _P, _P, C.c_int, _P]),
    'stv_vpad': (C.c_int, [C.POINTER(ConvGeom), _P, _P, C.c_int, _P, _P]),
    'stv_grad_pull': (C.c_int, [C.c_int]*4 + [_P] + [C.c_int]*4 + [_P, C.c_int, _P]),
    'stv_act_bwd': (C.c_int, [C.c_longlong, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    'stv_colsum': (C.c_int, [C.c_longlong, C.c_int, C.c_longlong, _P, _P, _P]),
    'stv_resample_bilinear': (C.c_int, [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _P, _P, _P]),
    'stv_mean_std_workspace_bytes': (C.c_size_t, [C.c_int]),
    'stv_mean_std': (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    'stv_ls_tail': (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'stv_rowscale': (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P]),
    'stv_bn_workspace_bytes': (C.c_size_t, [C.c_int]),
    'stv_bn_fwd': (C.c_int, [C.c_longlong, C.c_int, _P, _P, _P, _P, C.c_int, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'stv_bn_bwd': (C.c_int, [C.c_longlong, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P, C.c_size_t, _P]),
    'stv_maxpool3x3s2_fwd': (C.c_int, [C.c_int]*4 + [_P, _P, _P, _P]),
    'stv_maxpool3x3s2_bwd': (C.c_int, [C.c_int]*4 + [_P, _P, _P, _P]),
    'stv_head3x3_fwd': (C.c_int, [C.c_int]*4 + [_P, _P, _P, C.c_int, _P, _P]),
    'stv_head3x3_bwd': (C.c_int, [C.c_int]*4 + [_P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    'stv_inv4x4': (C.c_int, [C.c_int, _P, _P, _P]),
    'stv_inv4x4_bwd': (C.c_int, [C.c_int, _P, _P, _P, _P]),
    'stv_adamw_step': (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.c_size_t] + [C.c_float]*6 + [C.c_int, _P]),
}


def exported_symbols() -> list[str]:
    """Every entry point include/stv.h declares (checked by the CPU test-suite against the built library)."""
    return list(_SIGNATURES)


def lib() -> C.CDLL:
    """Load libstv.so once. Raises if it has not been built (`python -m slowtv_monodepth_b200._build`)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.is_file():
            raise StvError(f'{LIB_PATH} is missing: build it with `python -m slowtv_monodepth_b200._build` '
                           f'(there is no CPU/PyTorch fallback for the CUDA path).')
        h = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the symbol is not exported.
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


# The product path is CUDA-only: modules handed host (CPU) tensors raise. The package contains NO host arithmetic. The CPU
# test-suite checks parameter naming / wiring / the multi-process optimiser logic on the CPU with reference implementations that
# live in tests/host_ref.py and are registered here by the test fixtures (never by product code).
_HOST_HOOKS: dict = {}


def host_path(obj, *args, what: str | None = None, **kwargs):
    """Dispatch a host-tensor call to the hook a TEST registered for type(obj); raise in production (no hook is ever registered)."""
    fn = _HOST_HOOKS.get(type(obj))
    if fn is None:
        raise StvError(f'{what or type(obj).__name__}: got host (CPU) tensors. The product path is CUDA-only (libstv kernels); there '
                       f'is no CPU fallback.')
    return fn(obj, *args, **kwargs)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().stv_last_error().decode()
        if rc == 1: raise ValueError(f'{what}: {msg}')  # Reference convention: ValueError for bad shapes/args.
        raise StvError(f'{what} failed (code {rc}): {msg}')


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def ptr_array(ts) -> C.Array:
    return (C.c_void_p*len(ts))(*[t.data_ptr() for t in ts])


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*ts: torch.Tensor | None, what: str = 'stv') -> None:
    for t in ts:
        if t is None: continue
        if not t.is_cuda:
            raise StvError(f'{what}: expected CUDA tensors, got a {t.device} tensor (there is no CPU fallback).')
        if t.dtype not in (torch.float32, torch.uint8, torch.int64):
            raise ValueError(f'{what}: expected float32 tensors, got {t.dtype}.')


def launch_count() -> int:
    return int(lib().stv_launch_count())
