"""Autograd bindings of the libstv kernels (the host side of the C ABI in include/stv.h).

Each `torch.autograd.Function` enqueues hand-written sm_100a kernels on the current CUDA stream through ctypes; torch is
used only for device memory, streams and the autograd graph around the kernels.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib as L

__all__ = ['photo_loss', 'photo_error', 'smooth_loss', 'disp_to_depth', 'view_synth', 'adamw_step_', 'dwconv7', 'layer_norm', 'gemm_tf32']


# Optional per-call device timing of the fused loss kernels (bench.py's roofline figures): when enabled, a pair of CUDA
# events brackets each library call on the launching stream; `kernel_timings()` resolves them after a synchronize.
_TIMING: dict[str, list] | None = None
_TIMING_DETAIL = False


def enable_kernel_timing(on: bool = True, detail: bool = False) -> None:
    """detail=True keys every call by entry point AND problem shape (tools/gemm_breakdown.py)."""
    global _TIMING, _TIMING_DETAIL
    _TIMING = {} if on else None
    _TIMING_DETAIL = bool(on and detail)


class _timed:
    def __init__(self, name: str, detail=None): self.name, self.detail = name, detail

    def __enter__(self):
        if _TIMING is not None:
            if _TIMING_DETAIL and self.detail is not None:
                d = self.detail
                if isinstance(d, L.ConvGeom): d = ' '.join(f'{k}={getattr(d, k)}' for k, _ in d._fields_)
                self.name = f'{self.name} {d}'
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *a):
        if _TIMING is not None:
            self.ev[1].record()
            _TIMING.setdefault(self.name, []).append(self.ev)


def kernel_timings() -> dict[str, list[float]]:
    """Milliseconds per recorded call, keyed by entry point. Call after torch.cuda.synchronize()."""
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in (_TIMING or {}).items()}


def reset_kernel_timings() -> None:
    if _TIMING is not None: _TIMING.clear()


# Gradient sink: when a parameter's `.grad` already exists as a dense fp32 buffer in the kernel's own layout (FlatAdamW exposes
# every gradient as a view of one flat buffer, zeroed once per step), the weight-/bias-gradient kernels accumulate straight
# into it (their epilogues are atomic accumulations anyway) and autograd receives None for that input — instead of
# zero-filling a temporary, accumulating into it and letting AccumulateGrad add it to `.grad` (two extra launches and three
# extra passes over every parameter-sized tensor per step). Off by default; FlatAdamW switches it on.
GRAD_SINK = False


def grad_sink(on: bool = True) -> None:
    global GRAD_SINK
    GRAD_SINK = bool(on)


def _sink(p: Tensor | None, phys=None) -> Tensor | None:
    """`p.grad` viewed in the layout `phys` maps p to, when that view is the contiguous memory itself; else None."""
    if not GRAD_SINK or p is None or not p.is_leaf or p.grad is None or not p.grad.is_cuda or p.grad.dtype != torch.float32: return None
    g = p.grad if phys is None else phys(p.grad)
    return g if g.is_contiguous() else None


# Auxiliary streams. (1) Weight-gradient side streams: in backward, a layer's weight gradient is off the critical path (nothing
# downstream reads it before the optimiser) while its data gradient feeds the next layer. With the gradient sink on, the
# weight-gradient kernels have no autograd output at all, so they are enqueued on a second stream that forks from the stream of
# the backward node and joins the caller's stream once, when backward ends (autograd engine callback): the GPU co-schedules their
# CTAs with the data-gradient kernels of the following layers, filling the partial waves and the prologue / epilogue bubbles of
# these ~10 GFLOP products. (2) Branch streams (`branch_stream`): the pose network runs beside the depth network, forward AND
# backward (autograd replays every node on the stream of its forward). Captured CUDA graphs record the forks / joins as graph
# edges. Tensors read across streams are `record_stream`ed (the caching allocator then defers their reuse; during capture until
# the capture ends).
import os as _os0
WGRAD_STREAM = not bool(_os0.environ.get('STV_WGRAD_STREAM_OFF'))     # developer switches for A/B runs
BRANCH_STREAMS = not bool(_os0.environ.get('STV_BRANCH_STREAM_OFF'))
_SIDE_STREAMS: dict[tuple, 'torch.cuda.Stream'] = {}    # (device, origin stream handle) -> weight-gradient stream
_BRANCH_STREAMS: dict[tuple, 'torch.cuda.Stream'] = {}  # (device, name) -> branch stream
_DIRTY: set = set()                                     # streams with work not yet joined into the caller's stream
_JOIN_QUEUED = [False]


def _join_all() -> None:
    """The CURRENT stream waits for every auxiliary stream that has un-joined work (no host synchronisation)."""
    _JOIN_QUEUED[0] = False
    cur = torch.cuda.current_stream()
    for st in list(_DIRTY):
        if st.device == cur.device and st != cur:   # (a branch stream joining others stays pending itself)
            cur.wait_stream(st)
            _DIRTY.discard(st)


def join_side_stream(device=None) -> None:
    """Make the current stream wait for the auxiliary streams (weight gradients, branches): call before reading gradients from a
    stream-ordered consumer outside autograd's own end-of-backward join (the optimiser and the all-reduce do)."""
    if _DIRTY: _join_all()


def _queue_join() -> bool:
    """Ask the autograd engine to join the auxiliary streams into the caller's stream when the running backward pass ends."""
    if _JOIN_QUEUED[0]: return True
    try: torch.autograd.Variable._execution_engine.queue_callback(_join_all)
    except RuntimeError: return False   # not inside a backward pass
    _JOIN_QUEUED[0] = True
    return True


class _side_stream:
    """`with _side_stream(t1, t2, ...):` — enqueue the enclosed libstv calls on the weight-gradient stream of the current stream
    (no-op when disabled or outside a backward pass). t_i: tensors the enclosed kernels read or write."""
    def __init__(self, *tensors): self.tensors, self.cm = tensors, None

    def __enter__(self):
        if not WGRAD_STREAM or not GRAD_SINK: return self
        t0 = next((t for t in self.tensors if t is not None), None)
        if t0 is None or not t0.is_cuda or not _queue_join(): return self
        dev = t0.device.index
        cur = torch.cuda.current_stream(dev)
        key = (dev, cur.cuda_stream)
        side = _SIDE_STREAMS.get(key)
        if side is None: side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        for t in self.tensors:
            if t is not None: t.record_stream(side)
        _DIRTY.add(side)
        self.cm = torch.cuda.stream(side)
        self.cm.__enter__()
        return self

    def __exit__(self, *a):
        if self.cm is not None: self.cm.__exit__(*a)
        return False


class branch_stream:
    """`with branch_stream('pose', inputs) as br: ...; br.outputs(tensors)` — run an independent branch of the forward pass on its
    own stream, forked from the current one; `join()` (called by the user after the concurrent work has been enqueued) makes the
    current stream wait for it. Backward replays the branch on the same stream; its tail is joined when backward ends."""
    def __init__(self, name: str, inputs=()):
        self.name, self.inputs, self.cm, self.st, self.cur = name, [t for t in inputs if torch.is_tensor(t) and t.is_cuda], None, None, None

    def __enter__(self):
        if not BRANCH_STREAMS or not self.inputs: return self
        dev = self.inputs[0].device.index
        self.cur = torch.cuda.current_stream(dev)
        key = (dev, self.name)
        st = _BRANCH_STREAMS.get(key)
        if st is None: st = _BRANCH_STREAMS[key] = torch.cuda.Stream(device=dev)
        st.wait_stream(self.cur)
        for t in self.inputs: t.record_stream(st)
        self.st = st
        self.cm = torch.cuda.stream(st)
        self.cm.__enter__()
        return self

    def __exit__(self, *a):
        if self.cm is not None: self.cm.__exit__(*a)
        return False

    def join(self, outputs=()) -> None:
        """The forking stream waits for the branch; `outputs` (tensors produced on the branch, consumed on the forking stream) are
        recorded there, and the first one that takes part in autograd arms the end-of-backward join of the branch stream."""
        if self.st is None: return
        self.cur.wait_stream(self.st)
        armed = False
        for t in outputs:
            if not (torch.is_tensor(t) and t.is_cuda): continue
            t.record_stream(self.cur)
            if t.requires_grad and not armed:
                st = self.st
                t.register_hook(lambda g, st=st: (_DIRTY.add(st), _queue_join(), g)[2])
                armed = True


def _f32c(t: Tensor | None) -> Tensor | None:
    if t is None: return None
    if t.dtype != torch.float32: raise ValueError(f'Expected float32, got {t.dtype}.')
    return t.contiguous()


def _ws(nbytes: int, device) -> Tensor:
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------------------------------------------------
# Fused view-synthesis photometric loss
# ---------------------------------------------------------------------------------------------------------------------
# Developer switches (parity / property tests): route min-reprojection through the two-pass kernels as well; keep a reference to
# the single-pass kernel's full-resolution unit-gradient maps of the most recent call.
PHOTO_FORCE_TWO_PASS = False
# Reproducible mode (STV_DETERMINISTIC=1, read by libstv too: include/stv.h): bit-identical gradients from run to run, several times slower.
DETERMINISTIC = _os0.environ.get('STV_DETERMINISTIC', '') == '1'
KEEP_UNIT_GRADS = False
LAST_UNIT_GRADS = None


class _PhotoLoss(torch.autograd.Function):
    """Two-pass formulation (stv_photo_fwd + stv_photo_bwd; the backward re-warps a halo-2 tile and rebuilds the SSIM sums):
    any reduction (mean / min), any number of support frames. The default configuration goes through `_PhotoFused`."""
    @staticmethod
    def forward(ctx, cfg: L.PhotoCfg, want_warp: bool, tgt, supp, T, K, Kinv, noise, noise_step, *depths):
        L.require_cuda(tgt, supp, T, K, Kinv, noise, noise_step, *depths, what='photo_loss')
        lib, dev = L.lib(), tgt.device
        b, n, S, H, W = cfg.b, cfg.n, cfg.S, cfg.H, cfg.W
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            sel = torch.empty((S, b, H, W), dtype=torch.uint8, device=dev)
            warp0 = torch.empty((n, b, 3, H, W), dtype=torch.float32, device=dev) if want_warp else None
            nws = lib.stv_photo_workspace_bytes(C.byref(cfg))
            ws = _ws(nws, dev)
            with _timed('stv_photo_fwd'):
                L.check(lib.stv_photo_fwd(C.byref(cfg), L.ptr_array(depths), L.ptr(tgt), L.ptr(supp), L.ptr(T), L.ptr(K),
                                          L.ptr(Kinv), L.ptr(noise), L.ptr(noise_step), L.ptr(loss), L.ptr(sel), L.ptr(warp0), L.ptr(ws),
                                          ws.numel(), L.stream()), 'stv_photo_fwd')
        ctx.cfg, ctx.nws = cfg, nws
        ctx.save_for_backward(tgt, supp, T, K, Kinv, sel, *depths)
        ctx.mark_non_differentiable(sel)
        if warp0 is None: warp0 = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(warp0)
        return loss, sel, warp0

    @staticmethod
    def backward(ctx, g_loss, _g_sel, _g_warp):
        tgt, supp, T, K, Kinv, sel, *depths = ctx.saved_tensors
        cfg, lib, dev = ctx.cfg, L.lib(), tgt.device
        need_d = any(ctx.needs_input_grad[9:])
        need_T, need_K, need_Ki = ctx.needs_input_grad[4], ctx.needs_input_grad[5], ctx.needs_input_grad[6]
        if not (need_d or need_T or need_K or need_Ki): return (None,)*(9 + len(depths))
        with torch.cuda.device(dev):
            g_loss = g_loss.to(torch.float32).contiguous()
            g_depths = [torch.empty_like(d) for d in depths]
            gT = torch.empty_like(T)
            want_k = need_K or need_Ki
            gK = torch.empty_like(K) if want_k else None
            gKi = torch.empty_like(Kinv) if want_k else None
            ws = _ws(ctx.nws, dev)
            with _timed('stv_photo_bwd'):
                L.check(lib.stv_photo_bwd(C.byref(cfg), L.ptr_array(depths), L.ptr(tgt), L.ptr(supp), L.ptr(T), L.ptr(K),
                                          L.ptr(Kinv), L.ptr(sel), L.ptr(g_loss), L.ptr_array(g_depths), L.ptr(gT), L.ptr(gK),
                                          L.ptr(gKi), L.ptr(ws), ws.numel(), L.stream()), 'stv_photo_bwd')
        return (None, None, None, None, gT if need_T else None, gK if need_K else None, gKi if need_Ki else None, None, None,
                *[g if ctx.needs_input_grad[9 + j] else None for j, g in enumerate(g_depths)])


# Texture views of the support frames for the 2x2 gathers: handles are created by the library on request and OWNED HERE (the
# library keeps no state). Keyed by (device, address, shape): torch's caching allocator hands the same buffers back every step
# and CUDA-graph runners use static input buffers, so steady state creates none; least-recently-used views are destroyed.
import os as _os
_NO_TEX = bool(_os.environ.get('STV_NO_TEX'))   # developer switch: plain-load variant of the gathers
_TEX: dict[tuple, int] = {}
_TEX_MAX = 64


def _tex_handle(supp: Tensor) -> int:
    if _NO_TEX: return 0
    rows, W = supp.numel()//supp.shape[-1], supp.shape[-1]
    key = (supp.device.index, supp.data_ptr(), rows, W)
    h = _TEX.pop(key, None)
    if h is None:
        out = C.c_ulonglong(0)
        with torch.cuda.device(supp.device):
            L.check(L.lib().stv_tex_create(L.ptr(supp), rows, W, C.byref(out)), 'stv_tex_create')
        h = int(out.value)
        if len(_TEX) >= _TEX_MAX:
            old = next(iter(_TEX))
            L.lib().stv_tex_destroy(_TEX.pop(old))
    _TEX[key] = h  # most recently used last
    return h


class _PhotoFused(torch.autograd.Function):
    """Single-pass kernel (stv_photo_fused_fwd): the loss and its UNIT gradients in one sweep; the backward only scales them
    (and pulls the per-pixel maps back through the bilinear up-sampling when the kernel consumed low-resolution disparities)."""
    @staticmethod
    def forward(ctx, cfg: L.PhotoCfg, src: L.PhotoSrc, want_warp: bool, tgt, supp, T, K, Kinv, noise, noise_step, *maps):
        L.require_cuda(tgt, supp, T, K, Kinv, noise, noise_step, *maps, what='photo_loss')
        lib, dev = L.lib(), tgt.device
        b, n, S, H, W = cfg.b, cfg.n, cfg.S, cfg.H, cfg.W
        grad = any(ctx.needs_input_grad[5:8] + ctx.needs_input_grad[10:])
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            sel = torch.empty((S, b, H, W), dtype=torch.uint8, device=dev)
            warp0 = torch.empty((n, b, 3, H, W), dtype=torch.float32, device=dev) if want_warp else None
            g_unit = [torch.empty((b, 1, H, W), dtype=torch.float32, device=dev) for _ in range(S)] if grad else None
            gpart = _ws(lib.stv_photo_fused_partial_bytes(C.byref(cfg)), dev) if grad else None
            ws = _ws(lib.stv_photo_fused_workspace_bytes(C.byref(cfg)), dev)
            with _timed('stv_photo_fwd'):
                L.check(lib.stv_photo_fused_fwd(C.byref(cfg), C.byref(src), L.ptr_array(maps), L.ptr(tgt), L.ptr(supp), _tex_handle(supp),
                                                L.ptr(T), L.ptr(K), L.ptr(Kinv), L.ptr(noise), L.ptr(noise_step), L.ptr(loss), L.ptr(sel),
                                                L.ptr(warp0), L.ptr_array(g_unit) if grad else None, L.ptr(gpart), L.ptr(ws), ws.numel(),
                                                L.stream()), 'stv_photo_fused_fwd')
        ctx.cfg, ctx.src, ctx.shapes = cfg, src, [m.shape for m in maps]
        if KEEP_UNIT_GRADS:
            global LAST_UNIT_GRADS
            LAST_UNIT_GRADS = g_unit
        ctx.save_for_backward(T, Kinv, gpart, *(g_unit or []))
        ctx.mark_non_differentiable(sel)
        if warp0 is None: warp0 = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(warp0)
        return loss, sel, warp0

    @staticmethod
    def backward(ctx, g_loss, _g_sel, _g_warp):
        T, Kinv, gpart, *g_unit = ctx.saved_tensors
        cfg, src, lib, dev = ctx.cfg, ctx.src, L.lib(), T.device
        nmap = len(ctx.shapes)
        need_m = any(ctx.needs_input_grad[10:])
        need_T, need_K, need_Ki = ctx.needs_input_grad[5], ctx.needs_input_grad[6], ctx.needs_input_grad[7]
        if not (need_m or need_T or need_K or need_Ki) or gpart is None: return (None,)*(10 + nmap)
        with torch.cuda.device(dev):
            g_loss = g_loss.to(torch.float32).contiguous()
            g_maps = [torch.empty(sh, dtype=torch.float32, device=dev) for sh in ctx.shapes] if need_m else None
            want_p = need_T or need_K or need_Ki
            gT = torch.empty_like(T) if want_p else None
            gK = torch.empty_like(Kinv) if (need_K or need_Ki) else None
            gKi = torch.empty_like(Kinv) if (need_K or need_Ki) else None
            ws = _ws(lib.stv_photo_fused_bwd_workspace_bytes(C.byref(cfg), C.byref(src)), dev)
            with _timed('stv_photo_bwd'):
                L.check(lib.stv_photo_fused_bwd(C.byref(cfg), C.byref(src), L.ptr(g_loss), L.ptr_array(g_unit), L.ptr(gpart), L.ptr(T), L.ptr(Kinv),
                                                L.ptr_array(g_maps) if need_m else None, L.ptr(gT), L.ptr(gK), L.ptr(gKi), L.ptr(ws),
                                                ws.numel(), L.stream()), 'stv_photo_fused_bwd')
        return (None, None, None, None, None, gT if need_T else None, gK if need_K else None, gKi if need_Ki else None, None, None,
                *[g if ctx.needs_input_grad[10 + j] else None for j, g in enumerate(g_maps or [None]*nmap)])


FUSED_MAX_SUPPORT = 4


def _photo_cfg(b, n, S, H, W, loss_name, use_min, use_automask, noise_seed) -> L.PhotoCfg:
    if loss_name == 'ssim': w_ssim, w_l1 = 0.85, 0.15  # PhotoError(weight_ssim=0.85), src/losses/reconstruction.py:38
    elif loss_name == 'l1': w_ssim, w_l1 = 0.0, 1.0    # DenseL1Error, reconstruction.py:39
    else: raise ValueError(f'The fused photometric loss supports loss_name in {{ssim, l1}}, got "{loss_name}".')
    return L.PhotoCfg(b=b, n=n, S=S, H=H, W=W, w_ssim=w_ssim, w_l1=w_l1, use_min=int(use_min),
                      use_automask=int(use_automask), noise_seed=int(noise_seed), depth_stride_s=0)


def photo_loss(depths: list[Tensor], tgt: Tensor, supp: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None, *,
               loss_name: str = 'ssim', use_min: bool = True, use_automask: bool = True, noise: Tensor | None = None,
               noise_seed: int = 0, noise_step: Tensor | None = None, want_warp: bool = False,
               disp_size: tuple[int, int] | None = None, min_depth: float | None = None, max_depth: float | None = None):
    """Fused warp + photometric loss over all scales and support frames.

    depths: S x (b,1,H,W) up-sampled depth maps — or, with `disp_size=(H, W)`, S x (b,1,h_s,w_s) sigmoid disparities straight
    from the network: the bilinear up-sampling (ops.interpolate_like) and to_scaled / to_inv with `min_depth` / `max_depth` are then
    done inside the kernel and the gradient comes back at the disparities' own resolution (src/core/trainer.py:320-321 fused away).
    tgt (b,3,H,W); supp (n,b,3,H,W); T (n,b,4,4); K (b,4,4); K_inv (b,4,4) or None (= K^-1, differentiable, as
    `ViewSynth.forward` does with `K.inverse()`, src/tools/geometry.py:383).
    noise: None or (S*b,1,H,W) explicit tie-break noise (parity tests); otherwise `noise_seed != 0` draws it in-kernel with the
    effective seed noise_seed + noise_step[0]; `noise_step` (one int64 on the device, optional) is advanced by the call itself, so
    every call — and every replay of a captured CUDA graph — draws fresh noise (torch.randn_like, reconstruction.py:72).
    -> loss (), sel (S,b,H,W) uint8, warp0 (n,b,3,H,W) | None.
    """
    S = len(depths)
    b, _, H, W = tgt.shape
    n = supp.shape[0]
    if supp.shape != (n, b, 3, H, W): raise ValueError(f'Invalid support frames shape. ({tuple(supp.shape)} vs. {(n, b, 3, H, W)})')
    if T.shape != (n, b, 4, 4): raise ValueError(f'Invalid transforms shape. ({tuple(T.shape)} vs. {(n, b, 4, 4)})')
    if K.shape != (b, 4, 4): raise ValueError(f'Invalid intrinsics shape. ({tuple(K.shape)} vs. {(b, 4, 4)})')
    if S > L.MAX_SCALES: raise ValueError(f'At most {L.MAX_SCALES} scales are supported, got {S}.')
    from_disp = disp_size is not None
    if from_disp and tuple(disp_size) != (H, W): raise ValueError(f'disp_size {tuple(disp_size)} does not match the frames {(H, W)}.')
    for d in depths:
        if from_disp:
            if d.ndim != 4 or d.shape[:2] != (b, 1): raise ValueError(f'Invalid disparity shape. ({tuple(d.shape)} vs. {(b, 1, "h", "w")})')
        elif d.shape != (b, 1, H, W): raise ValueError(f'Invalid depth shape. ({tuple(d.shape)} vs. {(b, 1, H, W)})')
    if noise is not None and noise.numel() != S*b*H*W:
        raise ValueError(f'Invalid noise shape. ({tuple(noise.shape)} vs. {(S*b, 1, H, W)})')
    if noise_step is not None and (noise_step.dtype != torch.int64 or noise_step.numel() != 1):
        raise ValueError('noise_step must be a single int64 element on the device.')
    if K_inv is None: K_inv = inv4x4(K)
    cfg = _photo_cfg(b, n, S, H, W, loss_name, use_min, use_automask, noise_seed)
    args = (_f32c(tgt), _f32c(supp), _f32c(T), _f32c(K), _f32c(K_inv), _f32c(noise), noise_step)
    if use_min and n <= FUSED_MAX_SUPPORT and not PHOTO_FORCE_TWO_PASS:
        src = L.PhotoSrc(mode=1 if from_disp else 0, min_depth=float(min_depth or 0), max_depth=float(max_depth or 0))
        for j, d in enumerate(depths): src.h[j], src.w[j] = (d.shape[2], d.shape[3]) if from_disp else (H, W)
        loss, sel, warp0 = _PhotoFused.apply(cfg, src, want_warp, *args, *[_f32c(d) for d in depths])
    else:
        if from_disp: depths = [disp_to_depth(d, (H, W), min_depth, max_depth)[1] for d in depths]
        loss, sel, warp0 = _PhotoLoss.apply(cfg, want_warp, *args, *[_f32c(d) for d in depths])
    return loss, sel, (warp0 if want_warp else None)


def photo_error(pred: Tensor, target: Tensor, *, loss_name: str = 'ssim', use_min: bool = True) -> Tensor:
    """`ReconstructionLoss.compute_photo` without gradient: pred (n,b,3,H,W) | (b,3,H,W), target (b,3,H,W) -> (b,1,H,W)."""
    if pred.ndim == 4: pred = pred[None]
    n, b, _, H, W = pred.shape
    L.require_cuda(pred, target, what='photo_error')
    cfg = _photo_cfg(b, n, 1, H, W, loss_name, use_min, False, 0)
    pred, target = _f32c(pred.detach()), _f32c(target.detach())
    with torch.cuda.device(pred.device):
        err = torch.empty((b, 1, H, W), dtype=torch.float32, device=pred.device)
        L.check(L.lib().stv_photo_error(C.byref(cfg), L.ptr(pred), L.ptr(target), L.ptr(err), L.stream()), 'stv_photo_error')
    return err


class _ReconLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: L.PhotoCfg, pred, tgt, source, noise, noise_step):
        L.require_cuda(pred, tgt, source, noise, noise_step, what='recon_loss')
        lib, dev = L.lib(), pred.device
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            sel = torch.empty((cfg.b, cfg.H, cfg.W), dtype=torch.uint8, device=dev)
            ws = _ws(lib.stv_recon_workspace_bytes(C.byref(cfg)), dev)
            L.check(lib.stv_recon_fwd(C.byref(cfg), L.ptr(pred), L.ptr(tgt), L.ptr(source), L.ptr(noise), L.ptr(noise_step), L.ptr(loss),
                                      L.ptr(sel), None, L.ptr(ws), ws.numel(), L.stream()), 'stv_recon_fwd')
        ctx.cfg = cfg
        ctx.save_for_backward(pred, tgt, sel)
        ctx.mark_non_differentiable(sel)
        return loss, sel

    @staticmethod
    def backward(ctx, g_loss, _g_sel):
        pred, tgt, sel = ctx.saved_tensors
        if not ctx.needs_input_grad[1]: return (None,)*6
        with torch.cuda.device(pred.device):
            g = torch.empty_like(pred)
            L.check(L.lib().stv_recon_bwd(C.byref(ctx.cfg), L.ptr(pred), L.ptr(tgt), L.ptr(sel), L.ptr(g_loss.to(torch.float32).contiguous()),
                                          L.ptr(g), L.stream()), 'stv_recon_bwd')
        return None, g, None, None, None, None


def recon_loss(pred: Tensor, target: Tensor, source: Tensor | None = None, *, loss_name: str = 'ssim', use_min: bool = False,
               use_automask: bool = False, noise: Tensor | None = None, noise_seed: int = 0, noise_step: Tensor | None = None):
    """`ReconstructionLoss.forward` on already warped frames (src/losses/reconstruction.py:98-126), differentiable in `pred`.
    pred (n,b,3,H,W) | (b,3,H,W); target (b,3,H,W); source like pred (needed when use_automask) -> loss (), sel (b,H,W) uint8."""
    if pred.ndim == 4: pred = pred[None]
    if source is not None and source.ndim == 4: source = source[None]
    n, b, c, H, W = pred.shape
    if c != 3 or target.shape != (b, 3, H, W): raise ValueError(f'Invalid shapes. ({tuple(pred.shape)} vs. {tuple(target.shape)})')
    if use_automask and source is None: raise ValueError("Must provide the original 'source' images when automasking...")
    if source is not None and source.shape != pred.shape: raise ValueError(f'Invalid source shape. ({tuple(source.shape)} vs. {tuple(pred.shape)})')
    cfg = _photo_cfg(b, n, 1, H, W, loss_name, use_min, use_automask, noise_seed)
    return _ReconLoss.apply(cfg, _f32c(pred), _f32c(target.detach()), _f32c(None if source is None else source.detach()),
                            _f32c(noise), noise_step)


class _ReconLossEx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: L.ReconCfg, pred, tgt, source, mask, noise, noise_step):
        L.require_cuda(pred, tgt, source, mask, noise, noise_step, what='recon_loss_ex')
        lib, dev = L.lib(), pred.device
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            sel = torch.empty((cfg.b, cfg.H, cfg.W), dtype=torch.uint8, device=dev)
            ws = _ws(lib.stv_recon_ex_workspace_bytes(C.byref(cfg)), dev)
            L.check(lib.stv_recon_ex_fwd(C.byref(cfg), L.ptr(pred), L.ptr(tgt), L.ptr(source), L.ptr(mask), L.ptr(noise), L.ptr(noise_step),
                                         L.ptr(loss), L.ptr(sel), None, L.ptr(ws), ws.numel(), L.stream()), 'stv_recon_ex_fwd')
        ctx.cfg = cfg
        ctx.save_for_backward(pred, tgt, source, mask, sel)
        ctx.mark_non_differentiable(sel)
        return loss, sel

    @staticmethod
    def backward(ctx, g_loss, _g_sel):
        pred, tgt, source, mask, sel = ctx.saved_tensors
        want_pred, want_mask = ctx.needs_input_grad[1], mask is not None and ctx.needs_input_grad[4]
        if not (want_pred or want_mask): return (None,)*7
        with torch.cuda.device(pred.device):
            g_pred = torch.empty_like(pred) if want_pred else None
            g_mask = torch.empty_like(mask) if want_mask else None
            L.check(L.lib().stv_recon_ex_bwd(C.byref(ctx.cfg), L.ptr(pred), L.ptr(tgt), L.ptr(source), L.ptr(mask), L.ptr(sel),
                                             L.ptr(g_loss.to(torch.float32).contiguous()), L.ptr(g_pred), L.ptr(g_mask), L.stream()),
                    'stv_recon_ex_bwd')
        return None, g_pred, None, None, g_mask, None, None


def _recon_ex_args(pred, target, source, mask, loss_name, use_min, use_automask, mask_name, noise_seed):
    if loss_name not in L.RECON_LOSS: raise KeyError(f'loss_name="{loss_name}" (ssim | l1 | l2)')
    if mask_name not in L.RECON_MASK: raise ValueError(f'Invalid mask type: {mask_name}')
    if pred.ndim == 4: pred = pred[None]
    if source is not None and source.ndim == 4: source = source[None]
    n, b, c, H, W = pred.shape
    if target.shape != (b, c, H, W): raise ValueError(f'Invalid shapes. ({tuple(pred.shape)} vs. {tuple(target.shape)})')
    if use_automask and source is None: raise ValueError("Must provide the original 'source' images when automasking...")
    if source is not None and source.shape != pred.shape: raise ValueError(f'Invalid source shape. ({tuple(source.shape)} vs. {tuple(pred.shape)})')
    if mask_name and mask is None: raise ValueError("Must provide a 'mask' when masking...")
    if mask is not None and not mask_name: mask = None   # apply_mask ignores the tensor without a mask_name (reconstruction.py:55-56)
    if mask is not None and mask.shape == (b, 1, H, W) and n > 1: mask = mask.expand(b, n, H, W)   # broadcasts like err*mask
    if mask is not None and mask.shape != (b, n, H, W): raise ValueError(f'Invalid mask shape. ({tuple(mask.shape)} vs. {(b, n, H, W)})')
    cfg = L.ReconCfg(b=b, n=n, C=c, H=H, W=W, loss=L.RECON_LOSS[loss_name], use_min=int(use_min), use_automask=int(use_automask),
                     mask_mode=L.RECON_MASK[mask_name if mask is not None else None], noise_seed=int(noise_seed))
    return cfg, pred, source, mask


def recon_loss_ex(pred: Tensor, target: Tensor, source: Tensor | None = None, mask: Tensor | None = None, *, loss_name: str = 'ssim',
                  use_min: bool = False, use_automask: bool = False, mask_name: str | None = None, noise: Tensor | None = None,
                  noise_seed: int = 0, noise_step: Tensor | None = None):
    """The full `ReconstructionLoss.forward` contract (src/losses/reconstruction.py:98-126): any channel count, loss_name
    ssim | l1 | l2, explainability / uncertainty masks; differentiable in `pred` and `mask`.
    pred (n,b,C,H,W) | (b,C,H,W); target (b,C,H,W); source like pred; mask (b,n,H,W) -> loss (), sel (b,H,W) uint8 (bit 7: automasked)."""
    cfg, pred, source, mask = _recon_ex_args(pred, target, source, mask, loss_name, use_min, use_automask, mask_name, noise_seed)
    return _ReconLossEx.apply(cfg, _f32c(pred), _f32c(target.detach()), _f32c(None if source is None else source.detach()), _f32c(mask),
                              _f32c(noise), noise_step)


def photo_error_ex(pred: Tensor, target: Tensor, mask: Tensor | None = None, *, loss_name: str = 'ssim', use_min: bool = True,
                   mask_name: str | None = None) -> Tensor:
    """`ReconstructionLoss.compute_photo` (reconstruction.py:79-96) for any channel count / loss / mask; forward only -> (b,1,H,W)."""
    cfg, pred, _, mask = _recon_ex_args(pred, target, None, mask, loss_name, use_min, False, mask_name, 0)
    L.require_cuda(pred, target, mask, what='photo_error_ex')
    pred, target, mask = _f32c(pred.detach()), _f32c(target.detach()), _f32c(None if mask is None else mask.detach())
    lib, dev = L.lib(), pred.device
    with torch.cuda.device(dev):
        err = torch.empty((cfg.b, 1, cfg.H, cfg.W), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        sel = torch.empty((cfg.b, cfg.H, cfg.W), dtype=torch.uint8, device=dev)
        ws = _ws(lib.stv_recon_ex_workspace_bytes(C.byref(cfg)), dev)
        L.check(lib.stv_recon_ex_fwd(C.byref(cfg), L.ptr(pred), L.ptr(target), None, L.ptr(mask), None, None, L.ptr(loss), L.ptr(sel),
                                     L.ptr(err), L.ptr(ws), ws.numel(), L.stream()), 'stv_recon_ex_fwd')
    return err


class _RegrLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, mask, loss_code: int, invert: bool):
        L.require_cuda(pred, target, mask, what='regr_loss')
        lib, dev, n = L.lib(), pred.device, pred.numel()
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            err = torch.empty_like(pred)
            ws = _ws(lib.stv_regr_workspace_bytes(), dev)
            L.check(lib.stv_regr_fwd(n, loss_code, int(invert), L.ptr(pred), L.ptr(target), L.ptr(mask), L.ptr(loss), L.ptr(err), L.ptr(ws),
                                     ws.numel(), L.stream()), 'stv_regr_fwd')
        ctx.args = (n, loss_code, int(invert))
        ctx.save_for_backward(pred, target, mask, ws)
        ctx.mark_non_differentiable(err)
        return loss, err

    @staticmethod
    def backward(ctx, g_loss, _g_err):
        pred, target, mask, ws = ctx.saved_tensors
        want_p, want_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_p or want_t): return (None,)*5
        with torch.cuda.device(pred.device):
            gp = torch.empty_like(pred) if want_p else None
            gt = torch.empty_like(target) if want_t else None
            L.check(L.lib().stv_regr_bwd(*ctx.args, L.ptr(pred), L.ptr(target), L.ptr(mask), L.ptr(g_loss.to(torch.float32).contiguous()),
                                         L.ptr(gp), L.ptr(gt), L.ptr(ws), ws.numel(), L.stream()), 'stv_regr_bwd')
        return gp, gt, None, None, None


def regr_loss(pred: Tensor, target: Tensor, mask: Tensor | None = None, *, loss_name: str = 'berhu', invert: bool = False):
    """`RegressionLoss.forward` (src/losses/regression.py:67-75): sum(mask e)/sum(mask) with e = l1 | log_l1 | berhu of (pred, target),
    optionally on to_inv of both; differentiable in `pred` and `target`. -> (loss (), err_regr like pred)."""
    if loss_name not in L.REGR_LOSS: raise KeyError(loss_name)
    if pred.shape != target.shape: raise ValueError(f'Non-matching shapes. ({tuple(pred.shape)} vs. {tuple(target.shape)})')
    if mask is not None:
        if mask.dtype != torch.float32: mask = mask.to(torch.float32)
        if mask.shape != pred.shape: mask = mask.expand_as(pred)
    return _RegrLoss.apply(_f32c(pred), _f32c(target), _f32c(None if mask is None else mask.detach()), L.REGR_LOSS[loss_name], bool(invert))


class _Inv4x4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A):
        L.require_cuda(A, what='inv4x4')
        with torch.cuda.device(A.device):
            B = torch.empty_like(A)
            L.check(L.lib().stv_inv4x4(A.numel()//16, L.ptr(A), L.ptr(B), L.stream()), 'stv_inv4x4')
        ctx.save_for_backward(B)
        return B

    @staticmethod
    def backward(ctx, gB):
        B, = ctx.saved_tensors
        gB = _f32c(gB)
        with torch.cuda.device(B.device):
            gA = torch.empty_like(B)
            L.check(L.lib().stv_inv4x4_bwd(B.numel()//16, L.ptr(B), L.ptr(gB), L.ptr(gA), L.stream()), 'stv_inv4x4_bwd')
        return gA


def inv4x4(A: Tensor) -> Tensor:
    """Differentiable inverse of (*, 4, 4) matrices — `K.inverse()` / `T.inverse()` without ATen's host-synchronising batched LU."""
    if A.shape[-2:] != (4, 4): raise ValueError(f'inv4x4: expected (*, 4, 4), got {tuple(A.shape)}')
    return _Inv4x4.apply(_f32c(A))


# ---------------------------------------------------------------------------------------------------------------------
# Edge-aware smoothness
# ---------------------------------------------------------------------------------------------------------------------
class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg: L.SmoothCfg, want_maps: bool, img, *disps):
        L.require_cuda(img, *disps, what='smooth_loss')
        lib, dev = L.lib(), img.device
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            dg = torch.empty_like(disps[0]) if want_maps else None
            ig = torch.empty_like(disps[0]) if want_maps else None
            ws = _ws(lib.stv_smooth_workspace_bytes(C.byref(cfg)), dev)
            with _timed('stv_smooth_fwd'):
                L.check(lib.stv_smooth_fwd(C.byref(cfg), L.ptr_array(disps), L.ptr(img), L.ptr(loss), L.ptr(dg), L.ptr(ig),
                                           L.ptr(ws), ws.numel(), L.stream()), 'stv_smooth_fwd')
        ctx.cfg = cfg
        ctx.save_for_backward(img, ws, *disps)
        if dg is None: dg = ig = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(dg, ig)
        return loss, dg, ig

    @staticmethod
    def backward(ctx, g_loss, _a, _b):
        img, ws, *disps = ctx.saved_tensors
        cfg, dev = ctx.cfg, img.device
        with torch.cuda.device(dev):
            g_loss = g_loss.to(torch.float32).contiguous()
            gds = [torch.empty_like(d) for d in disps]
            with _timed('stv_smooth_bwd'):
                L.check(L.lib().stv_smooth_bwd(C.byref(cfg), L.ptr_array(disps), L.ptr(img), L.ptr(g_loss), L.ptr_array(gds),
                                               L.ptr(ws), ws.numel(), L.stream()), 'stv_smooth_bwd')
        return (None, None, None, *gds)


def smooth_loss(disps: list[Tensor], img: Tensor, *, scales: list[int] | None = None, use_edges: bool = True,
                want_maps: bool = False):
    """All-scale edge-aware smoothness: mean_s(SmoothReg(disp_s, resize(img))/2**scale_s).
    disps: list of (b,1,h_s,w_s); img (b,3,H,W) -> loss (), disp_grad|None, image_grad|None (first scale)."""
    S = len(disps)
    if S > L.MAX_SCALES: raise ValueError(f'At most {L.MAX_SCALES} scales are supported, got {S}.')
    b, c, H, W = img.shape
    if c != 3: raise ValueError(f'Expected a 3-channel image, got {c} channels.')
    scales = list(range(S)) if scales is None else list(scales)
    cfg = L.SmoothCfg(b=b, S=S, H=H, W=W, use_edges=int(use_edges))
    for j, d in enumerate(disps):
        if d.ndim != 4 or d.shape[0] != b or d.shape[1] != 1: raise ValueError(f'Invalid disparity shape {tuple(d.shape)}.')
        cfg.h[j], cfg.w[j], cfg.scale_div[j] = d.shape[2], d.shape[3], float(2**scales[j])
    loss, dg, ig = _SmoothLoss.apply(cfg, want_maps, _f32c(img), *[_f32c(d) for d in disps])
    return loss, (dg if want_maps else None), (ig if want_maps else None)


# ---------------------------------------------------------------------------------------------------------------------
# Disparity -> upsampled depth
# ---------------------------------------------------------------------------------------------------------------------
class _DispToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, H: int, W: int, min_depth: float, max_depth: float):
        L.require_cuda(disp, what='disp_to_depth')
        b, _, h, w = disp.shape
        with torch.cuda.device(disp.device):
            disp_up = torch.empty((b, 1, H, W), dtype=torch.float32, device=disp.device)
            depth_up = torch.empty_like(disp_up)
            L.check(L.lib().stv_disp_to_depth_fwd(b, h, w, H, W, min_depth, max_depth, L.ptr(disp), L.ptr(disp_up),
                                                  L.ptr(depth_up), L.stream()), 'stv_disp_to_depth_fwd')
        ctx.save_for_backward(disp)
        ctx.args = (H, W, min_depth, max_depth)
        ctx.set_materialize_grads(False)
        return disp_up, depth_up

    @staticmethod
    def backward(ctx, g_disp_up, g_depth_up):
        if g_disp_up is None and g_depth_up is None: return None, None, None, None, None
        disp, = ctx.saved_tensors
        H, W, mn, mx = ctx.args
        b, _, h, w = disp.shape
        with torch.cuda.device(disp.device):
            g = torch.empty_like(disp)
            ws = _ws(L.lib().stv_disp_to_depth_bwd_workspace_bytes(b, h, w, H, W), disp.device)
            L.check(L.lib().stv_disp_to_depth_bwd(b, h, w, H, W, mn, mx, L.ptr(disp), L.ptr(_f32c(g_depth_up)),
                                                  L.ptr(_f32c(g_disp_up)), L.ptr(g), L.ptr(ws), ws.numel(), L.stream()), 'stv_disp_to_depth_bwd')
        return g, None, None, None, None


def disp_to_depth(disp: Tensor, size: tuple[int, int], min_depth: float | None, max_depth: float | None):
    """Bilinear upsample (align_corners=False) fused with to_scaled/to_inv: (b,1,h,w) -> disp_up, depth_up (b,1,H,W)."""
    if disp.ndim != 4 or disp.shape[1] != 1: raise ValueError(f'Invalid disparity shape {tuple(disp.shape)}.')
    if (min_depth or max_depth):
        if not min_depth or min_depth <= 0: raise ValueError(f'Min depth must be greater than 0. ({min_depth})')
        if max_depth and max_depth < min_depth: raise ValueError(f'Max depth must be greater than min. ({max_depth} vs. {min_depth})')
    disp_up, depth_up = _DispToDepth.apply(_f32c(disp), int(size[0]), int(size[1]), float(min_depth or 0), float(max_depth or 0))
    # Provenance tag: `handlers.image_recon` hands the network's disparity straight to the fused loss kernel when every depth map it
    # receives was produced here (same range), instead of reading these up-sampled copies and differentiating through them.
    depth_up._stv_src = (disp, (int(size[0]), int(size[1])), min_depth or None, max_depth or None)
    return disp_up, depth_up


# ---------------------------------------------------------------------------------------------------------------------
# Stand-alone ViewSynth
# ---------------------------------------------------------------------------------------------------------------------
class _ViewSynth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, depth, T, K, Kinv):
        L.require_cuda(inp, depth, T, K, Kinv, what='view_synth')
        B, Cc, H, W = inp.shape
        dev = inp.device
        with torch.cuda.device(dev):
            warp = torch.empty_like(inp)
            dwarp = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
            valid = torch.empty((B, 1, H, W), dtype=torch.uint8, device=dev)
            L.check(L.lib().stv_view_synth_fwd(B, Cc, H, W, L.ptr(inp), L.ptr(depth), L.ptr(T), L.ptr(K), L.ptr(Kinv),
                                               L.ptr(warp), L.ptr(dwarp), L.ptr(valid), L.stream()), 'stv_view_synth_fwd')
        ctx.save_for_backward(inp, depth, T, K, Kinv)
        valid = valid.bool()
        ctx.mark_non_differentiable(valid)
        return warp, dwarp, valid

    @staticmethod
    def backward(ctx, g_warp, g_dwarp, _g_valid):
        inp, depth, T, K, Kinv = ctx.saved_tensors
        B, Cc, H, W = inp.shape
        dev, lib = inp.device, L.lib()
        with torch.cuda.device(dev):
            g_depth, gT, gK, gKi = torch.empty_like(depth), torch.empty_like(T), torch.empty_like(K), torch.empty_like(Kinv)
            g_inp = torch.zeros_like(inp) if ctx.needs_input_grad[0] else None
            ws = _ws(lib.stv_view_synth_workspace_bytes(B, Cc, H, W), dev)
            L.check(lib.stv_view_synth_bwd(B, Cc, H, W, L.ptr(inp), L.ptr(depth), L.ptr(T), L.ptr(K), L.ptr(Kinv),
                                           L.ptr(_f32c(g_warp)), L.ptr(_f32c(g_dwarp)), L.ptr(g_depth), L.ptr(gT),
                                           L.ptr(gK), L.ptr(gKi), L.ptr(g_inp), L.ptr(ws), ws.numel(), L.stream()),
                    'stv_view_synth_bwd')
        return g_inp, g_depth, gT, gK, gKi


def view_synth(inp: Tensor, depth: Tensor, T: Tensor, K: Tensor, K_inv: Tensor | None = None):
    """`ViewSynth.forward` (src/tools/geometry.py:366-391): -> (input_warp, depth_warp, mask_valid)."""
    B, _, H, W = inp.shape
    if depth.shape != (B, 1, H, W): raise ValueError(f'Invalid depth shape. ({tuple(depth.shape)} vs. {(B, 1, H, W)})')
    if T.shape != (B, 4, 4) or K.shape != (B, 4, 4): raise ValueError(f'Invalid T/K shape. ({tuple(T.shape)}, {tuple(K.shape)})')
    if K_inv is None: K_inv = inv4x4(K)
    return _ViewSynth.apply(_f32c(inp), _f32c(depth), _f32c(T), _f32c(K), _f32c(K_inv))


# ---------------------------------------------------------------------------------------------------------------------
# ConvNeXt block pieces (channels-last)
# ---------------------------------------------------------------------------------------------------------------------
class _DwConv7(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, link=None):
        L.require_cuda(x, w, b, what='dwconv7')
        ctx.link = link
        N, H, W, Cc = x.shape
        with torch.cuda.device(x.device):
            y = torch.empty_like(x)
            L.check(L.lib().stv_dwconv7_fwd(N, H, W, Cc, L.ptr(x), L.ptr(w), L.ptr(b), None, L.ptr(y), 0, L.stream()), 'stv_dwconv7_fwd')
        ctx.save_for_backward(x, w)
        ctx.has_bias = b is not None
        ctx.w_sink, ctx.b_sink = _sink(w), _sink(b)
        if b is not None and (ctx.w_sink is None or ctx.b_sink is None): ctx.w_sink = ctx.b_sink = None  # one accumulate flag for both
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        N, H, W, Cc = x.shape
        gy = _f32c(gy)
        lib, dev = L.lib(), x.device
        gx = gw = gb = None
        with torch.cuda.device(dev):
            # ConvNeXt block: the gradient of the residual branch (left in `link` by _ConvNeXtMlp.backward, which always runs
            # first) is added by the data-gradient kernel itself instead of a separate autograd accumulation pass.
            g_res = ctx.link.pop('g_res', None) if ctx.link is not None else None
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                L.check(lib.stv_dwconv7_fwd(N, H, W, Cc, L.ptr(gy), L.ptr(w), None, L.ptr(g_res), L.ptr(gx), 1, L.stream()), 'stv_dwconv7_fwd(flip)')
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                sink = ctx.w_sink is not None
                gw = ctx.w_sink if sink else torch.empty_like(w)
                gb = (ctx.b_sink if sink else torch.empty(Cc, dtype=torch.float32, device=dev)) if ctx.has_bias else None
                ws = _ws(lib.stv_dwconv7_wgrad_workspace_bytes(N, H, W, Cc), dev)
                with (_side_stream(x, gy, ws) if sink else _side_stream()):
                    L.check(lib.stv_dwconv7_wgrad(N, H, W, Cc, L.ptr(x), L.ptr(gy), L.ptr(gw), L.ptr(gb), int(sink), L.ptr(ws), ws.numel(),
                                                  L.stream()), 'stv_dwconv7_wgrad')
                if sink: gw = gb = None  # accumulated in place into weight.grad / bias.grad
        return gx, gw, gb, None


def dwconv7(x: Tensor, weight: Tensor, bias: Tensor | None, link: dict | None = None) -> Tensor:
    """Depthwise 7x7 convolution, padding 3, on a channels-last (N,H,W,C) tensor; weight (C,1,7,7)."""
    if x.ndim != 4 or weight.shape != (x.shape[-1], 1, 7, 7): raise ValueError(f'dwconv7: bad shapes {tuple(x.shape)}, {tuple(weight.shape)}')
    return _DwConv7.apply(_f32c(x), _f32c(weight), _f32c(bias), link)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float):
        L.require_cuda(x, gamma, beta, what='layer_norm')
        Cc = x.shape[-1]
        P = x.numel()//Cc
        with torch.cuda.device(x.device):
            y = torch.empty_like(x)
            mean = torch.empty(P, dtype=torch.float32, device=x.device)
            rstd = torch.empty_like(mean)
            L.check(L.lib().stv_layernorm_fwd(P, Cc, L.ptr(x), L.ptr(gamma), L.ptr(beta), eps, L.ptr(y), L.ptr(mean), L.ptr(rstd),
                                              L.stream()), 'stv_layernorm_fwd')
        ctx.save_for_backward(x, mean, rstd, gamma)
        ctx.g_sink, ctx.b_sink = _sink(gamma), _sink(beta)
        if ctx.g_sink is None or ctx.b_sink is None: ctx.g_sink = ctx.b_sink = None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, gamma = ctx.saved_tensors
        Cc = x.shape[-1]
        P = x.numel()//Cc
        gy = _f32c(gy)
        lib, dev = L.lib(), x.device
        sink = ctx.g_sink is not None
        with torch.cuda.device(dev):
            gx = torch.empty_like(x)
            gg, gb = (ctx.g_sink, ctx.b_sink) if sink else (torch.empty_like(gamma), torch.empty_like(gamma))
            ws = _ws(lib.stv_layernorm_bwd_workspace_bytes(P, Cc), dev)
            L.check(lib.stv_layernorm_bwd(P, Cc, L.ptr(gy), L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(gx), L.ptr(gg),
                                          L.ptr(gb), int(sink), L.ptr(ws), ws.numel(), L.stream()), 'stv_layernorm_bwd')
        return (gx, None, None, None) if sink else (gx, gg, gb, None)


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm over the last axis of a contiguous tensor."""
    if gamma.shape != (x.shape[-1],): raise ValueError(f'layer_norm: bad shapes {tuple(x.shape)}, {tuple(gamma.shape)}')
    return _LayerNorm.apply(_f32c(x), _f32c(gamma), _f32c(beta), float(eps))


# ---------------------------------------------------------------------------------------------------------------------
# Tensor-core GEMM (tcgen05 kind::tf32) with fused epilogue
# ---------------------------------------------------------------------------------------------------------------------
def gemm_tf32(A: Tensor, B: Tensor, *, a_mn: bool = False, b_mn: bool = False, out: Tensor | None = None, bias: Tensor | None = None,
              act: str | None = None, aux: Tensor | None = None, gamma: Tensor | None = None, res: Tensor | None = None,
              dact: str | None = None, dact_src: Tensor | None = None, colsum: Tensor | None = None, accumulate: bool = False,
              split_k: int = 1) -> Tensor:
    """C[M,N] (+)= epilogue(sum_k A[m,k] B[n,k]) on the tcgen05 tensor cores (include/stv.h: stv_gemm_tf32). No autograd.

    A: (M,K) if not a_mn else (K,M);  B: (N,K) if not b_mn else (K,N); both 2-D with unit inner stride (row stride free).
    out / aux / res / dact_src: (M,N), same row stride. Returns `out` (allocated when None)."""
    L.require_cuda(A, B, out, bias, aux, gamma, res, dact_src, colsum, what='gemm_tf32')
    for t in (A, B):
        if t.ndim != 2 or t.stride(1) != 1:
            raise ValueError(f'gemm_tf32: operands must be 2-D with unit inner stride, got {tuple(t.shape)} / {t.stride()}')
    K, M = A.shape if a_mn else A.shape[::-1]
    Kb, N = B.shape if b_mn else B.shape[::-1]
    if K != Kb: raise ValueError(f'gemm_tf32: reduction sizes differ ({K} vs. {Kb})')
    dev = A.device
    with torch.cuda.device(dev):
        if out is None:
            if accumulate: raise ValueError('gemm_tf32: accumulate needs an existing `out`.')
            out = torch.empty((M, N), dtype=torch.float32, device=dev)
        if out.shape != (M, N) or out.stride(1) != 1: raise ValueError(f'gemm_tf32: bad output {tuple(out.shape)} / {out.stride()}')
        for t in (aux, res, dact_src):
            if t is not None and (t.shape != (M, N) or t.stride() != out.stride()):
                raise ValueError('gemm_tf32: aux / res / dact_src must match the output shape and strides.')
        for t in (bias, gamma, colsum):
            if t is not None and (t.shape != (N,) or not t.is_contiguous()): raise ValueError('gemm_tf32: bias / gamma / colsum must be contiguous (N,).')
        epi = L.GemmEpi(bias=L.ptr(bias), aux=L.ptr(aux), gamma=L.ptr(gamma), res=L.ptr(res), dact_src=L.ptr(dact_src), colsum=L.ptr(colsum),
                        act=L.ACT[act], dact=L.ACT[dact], accumulate=int(accumulate))
        with _timed('stv_gemm_tf32', (M, N, K, 'a_mn' if a_mn else '', 'b_mn' if b_mn else '', f'sk{split_k}')):
            L.check(L.lib().stv_gemm_tf32(M, N, K, L.ptr(A), A.stride(0), int(a_mn), L.ptr(B), B.stride(0), int(b_mn), L.ptr(out),
                                          out.stride(0), C.byref(epi), int(split_k), L.stream()), 'stv_gemm_tf32')
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Implicit-GEMM convolution (tcgen05) with fused upsample / concat / padding / bias / activation
# ---------------------------------------------------------------------------------------------------------------------
def _geom(src1: Tensor, src2: Tensor | None, w_phys: Tensor, up1: bool, stride: int, pad: int, reflect: bool) -> tuple[L.ConvGeom, int, int]:
    N, h1, w1, C1 = src1.shape
    H, W = (2*h1, 2*w1) if up1 else (h1, w1)
    C2 = 0 if src2 is None else src2.shape[3]
    if src2 is not None and src2.shape[:3] != (N, H, W): raise ValueError(f'conv2d_nhwc: skip tensor {tuple(src2.shape)} does not match {(N, H, W)}')
    Cout, R, S, Cin = w_phys.shape
    if Cin != C1 + C2: raise ValueError(f'conv2d_nhwc: filter expects {Cin} input channels, got {C1} + {C2}')
    g = L.ConvGeom(N=N, H=H, W=W, C1=C1, C2=C2, up1=int(up1), Cout=Cout, R=R, S=S, stride=stride, pad=pad, reflect=int(reflect))
    return g, (H + 2*pad - R)//stride + 1, (W + 2*pad - S)//stride + 1


def act_bwd(dA: Tensor, y: Tensor, act: str | None, dbias: Tensor | None = None) -> Tensor:
    """dZ = dA * act'(y) on (..., C) channels-last tensors; dbias (C) += column sums of dZ."""
    Cc = y.shape[-1]
    M = y.numel()//Cc
    dA = _f32c(dA)
    if act in (None, 'none'):  # identity: nothing to multiply — at most the bias gradient (column sums) is needed
        if dbias is not None:
            with torch.cuda.device(y.device):
                L.check(L.lib().stv_colsum(M, Cc, Cc, L.ptr(dA), L.ptr(dbias), L.stream()), 'stv_colsum')
        return dA
    with torch.cuda.device(y.device):
        dZ = torch.empty_like(y)
        L.check(L.lib().stv_act_bwd(M, Cc, L.ptr(dA), L.ptr(y), L.ACT[act], L.ptr(dZ), L.ptr(dbias), L.stream()), 'stv_act_bwd')
    return dZ


def colsum(x: Tensor) -> Tensor:
    """Column sums of a (..., C) contiguous tensor -> (C)."""
    Cc = x.shape[-1]
    with torch.cuda.device(x.device):
        out = torch.zeros(Cc, dtype=torch.float32, device=x.device)
        L.check(L.lib().stv_colsum(x.numel()//Cc, Cc, Cc, L.ptr(x), L.ptr(out), L.stream()), 'stv_colsum')
    return out


def grad_pull(dv: Tensor, shape: tuple[int, int, int, int], c_off: int, pad: int, pool: int) -> Tensor:
    """Gradient of one real source tensor (N,H,W,C) from the virtual-input gradient dv (N, H*pool+2pad, W*pool+2pad, Cs)."""
    N, H, W, Cc = shape
    with torch.cuda.device(dv.device):
        dst = torch.empty(shape, dtype=torch.float32, device=dv.device)
        L.check(L.lib().stv_grad_pull(N, H, W, Cc, L.ptr(dv), dv.shape[3], c_off, pad, pool, L.ptr(dst), 0, L.stream()), 'stv_grad_pull')
    return dst


class _Conv2dNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src1, src2, w, b, up1: bool, stride: int, pad: int, reflect: bool, act):
        L.require_cuda(src1, src2, w, b, what='conv2d_nhwc')
        w_phys = w.permute(0, 2, 3, 1).contiguous()  # (Cout,R,S,Cin): free for channels-last filters (the flat parameter buffer)
        g, P, Q = _geom(src1, src2, w_phys, up1, stride, pad, reflect)
        lib, dev = L.lib(), src1.device
        Cin = g.C1 + g.C2
        # The TMA im2col path needs one real, zero-padded tensor with channels % 32 == 0. A virtual input (reflection padding,
        # nearest x2 upsampling, skip concatenation) with such a channel count is materialised ONCE by stv_vpad (the reference
        # does it in three passes: interpolate, cat, pad) and convolved with pad 0; otherwise (3/6/16-channel layers) the
        # cp.async gather kernel resolves the virtual input on the fly.
        ctx.virt = None
        ctx.cin = Cin
        with torch.cuda.device(dev):
            virt = reflect or up1 or src2 is not None
            Cp = (Cin + 31)//32*32
            if (virt or Cp != Cin) and Cin >= 16:
                # Narrow layers (the 16-channel decoder level 0) are zero-padded to 32 channels on the way: the extra k-columns
                # multiply zero filter taps, and the TMA path is several times faster than the 16-byte gather.
                pv = pad if reflect else 0
                V = torch.empty((g.N, g.H + 2*pv, g.W + 2*pv, Cp), dtype=torch.float32, device=dev)
                L.check(lib.stv_vpad(C.byref(g), L.ptr(src1), L.ptr(src2), Cp, L.ptr(V), L.stream()), 'stv_vpad')
                ctx.virt = (tuple(src1.shape), None if src2 is None else tuple(src2.shape), g.C1, pv, 2 if up1 else 1)
                g = L.ConvGeom(N=g.N, H=g.H + 2*pv, W=g.W + 2*pv, C1=Cp, C2=0, up1=0, Cout=g.Cout, R=g.R, S=g.S, stride=stride,
                               pad=0 if reflect else pad, reflect=0)
                src1, src2 = V, None
                if Cp != Cin: w_phys = torch.nn.functional.pad(w_phys, (0, Cp - Cin))
            y = torch.empty((g.N, P, Q, g.Cout), dtype=torch.float32, device=dev)
            epi = L.GemmEpi(bias=L.ptr(b), act=L.ACT[act])
            with _timed('stv_conv_fprop', g):
                L.check(lib.stv_conv_fprop(C.byref(g), L.ptr(src1), L.ptr(src2), L.ptr(w_phys), L.ptr(y), C.byref(epi), L.stream()),
                        'stv_conv_fprop')
        ctx.save_for_backward(src1, src2, w_phys, y)
        ctx.g, ctx.act, ctx.has_bias = g, act, b is not None
        # gradient sinks (same memory layout as the kernel's outputs: un-padded filters only)
        ctx.w_sink = _sink(w, lambda t: t.permute(0, 2, 3, 1)) if (w_phys.shape[3] == w.shape[1] and g.Cout % 4 == 0) else None
        ctx.b_sink = _sink(b)
        return y

    @staticmethod
    def backward(ctx, dA):
        src1, src2, w_phys, y = ctx.saved_tensors
        g, act, lib, dev = ctx.g, ctx.act, L.lib(), y.device
        Cout, Cin = g.Cout, g.C1 + g.C2
        with torch.cuda.device(dev):
            db = None
            if ctx.has_bias: db = ctx.b_sink if ctx.b_sink is not None else torch.zeros(Cout, dtype=torch.float32, device=dev)
            dZ = act_bwd(dA, y, act, db)
            if ctx.b_sink is not None: db = None  # already accumulated into bias.grad
            wq = w_phys
            if Cout % 4:  # narrow heads (1-channel disparity): pad the output-channel axis to 4 for the TMA-fed operands
                cp = (-Cout) % 4
                dZ = torch.nn.functional.pad(dZ, (0, cp))
                wq = torch.nn.functional.pad(w_phys, (0, 0, 0, 0, 0, 0, 0, cp))
                g = L.ConvGeom.from_buffer_copy(g); g.Cout = Cout + cp
            dw = None
            if ctx.needs_input_grad[2]:
                dw = ctx.w_sink if ctx.w_sink is not None else torch.zeros_like(wq)
                with (_side_stream(src1, src2, dZ) if ctx.w_sink is not None else _side_stream()), _timed('stv_conv_wgrad', g):
                    L.check(lib.stv_conv_wgrad(C.byref(g), L.ptr(src1), L.ptr(src2), L.ptr(dZ), L.ptr(dw), 0, L.stream()), 'stv_conv_wgrad')
                dw = None if ctx.w_sink is not None else dw[:Cout, :, :, :ctx.cin].permute(0, 3, 1, 2)
            d1 = d2 = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                pd = g.pad if g.reflect else 0
                dv = torch.empty((g.N, g.H + 2*pd, g.W + 2*pd, Cin), dtype=torch.float32, device=dev)
                with _timed('stv_conv_dgrad', g):
                    L.check(lib.stv_conv_dgrad(C.byref(g), L.ptr(dZ), L.ptr(wq), L.ptr(dv), None, L.stream()), 'stv_conv_dgrad')
                if ctx.virt is not None:  # dv is the gradient of the materialised (padded) virtual input
                    shape1, shape2, c1, pv, pool = ctx.virt
                    if ctx.needs_input_grad[0]: d1 = grad_pull(dv, shape1, 0, pv, pool)
                    if shape2 is not None and ctx.needs_input_grad[1]: d2 = grad_pull(dv, shape2, c1, pv, 1)
                elif not g.reflect and not g.up1 and g.C2 == 0: d1 = dv
                else:
                    if ctx.needs_input_grad[0]: d1 = grad_pull(dv, tuple(src1.shape), 0, pd, 2 if g.up1 else 1)
                    if src2 is not None and ctx.needs_input_grad[1]: d2 = grad_pull(dv, tuple(src2.shape), g.C1, pd, 1)
        return d1, d2, dw, db, None, None, None, None, None


def conv2d_nhwc(src1: Tensor, w: Tensor, b: Tensor | None = None, *, src2: Tensor | None = None, up1: bool = False, stride: int = 1,
                pad: int = 0, reflect: bool = False, act: str | None = None) -> Tensor:
    """act(conv2d(cat(up2(src1) if up1 else src1, src2), w) + b) on channels-last tensors.

    src1 (N,h,w,C1), src2 (N,H,W,C2) | None: contiguous NHWC, C1 and C2 multiples of 4; w: the nn.Conv2d weight (Cout,Cin,R,S)
    (free when stored channels-last); -> (N,P,Q,Cout) contiguous NHWC. Differentiable in src1, src2, w, b."""
    if src1.ndim != 4 or w.ndim != 4: raise ValueError(f'conv2d_nhwc: expected 4-D tensors, got {tuple(src1.shape)}, {tuple(w.shape)}')
    if act not in (None, 'none', 'relu', 'elu', 'sigmoid'): raise ValueError(f'conv2d_nhwc: unsupported activation {act!r}')
    return _Conv2dNHWC.apply(_f32c(src1), _f32c(src2), w, _f32c(b), bool(up1), int(stride), int(pad), bool(reflect), act)


class _BatchNormNHWC(torch.autograd.Function):
    """Train-mode BatchNorm over the rows of a channels-last tensor, fused with the residual add and ReLU that follow."""
    @staticmethod
    def forward(ctx, x, gamma, beta, res, run_mean, run_var, relu: bool, eps: float, momentum: float):
        L.require_cuda(x, gamma, beta, res, what='batch_norm_nhwc')
        Cc = x.shape[-1]
        M = x.numel()//Cc
        lib, dev = L.lib(), x.device
        with torch.cuda.device(dev):
            y = torch.empty_like(x)
            mean = torch.empty(Cc, dtype=torch.float32, device=dev)
            rstd = torch.empty_like(mean)
            ws = _ws(lib.stv_bn_workspace_bytes(Cc), dev)
            L.check(lib.stv_bn_fwd(M, Cc, L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(res), int(relu), eps, momentum, L.ptr(y), L.ptr(mean),
                                   L.ptr(rstd), L.ptr(run_mean), L.ptr(run_var), L.ptr(ws), ws.numel(), L.stream()), 'stv_bn_fwd')
        ctx.save_for_backward(x, y, mean, rstd, gamma)
        ctx.relu, ctx.has_res = relu, res is not None
        ctx.g_sink, ctx.b_sink = _sink(gamma), _sink(beta)
        if ctx.g_sink is None or ctx.b_sink is None: ctx.g_sink = ctx.b_sink = None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, rstd, gamma = ctx.saved_tensors
        Cc = x.shape[-1]
        M = x.numel()//Cc
        lib, dev = L.lib(), x.device
        dy = _f32c(dy)
        with torch.cuda.device(dev):
            dx = torch.empty_like(x)
            dres = torch.empty_like(x) if ctx.has_res and ctx.relu else None
            sink = ctx.g_sink is not None
            dgamma, dbeta = (ctx.g_sink, ctx.b_sink) if sink else (torch.empty_like(gamma), torch.empty_like(gamma))
            ws = _ws(lib.stv_bn_workspace_bytes(Cc), dev)
            L.check(lib.stv_bn_bwd(M, Cc, L.ptr(dy), L.ptr(y), L.ptr(x), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), int(ctx.relu), L.ptr(dx),
                                   L.ptr(dres), L.ptr(dgamma), L.ptr(dbeta), int(sink), L.ptr(ws), ws.numel(), L.stream()), 'stv_bn_bwd')
        if ctx.has_res and not ctx.relu: dres = dy
        if sink: dgamma = dbeta = None
        return dx, dgamma, dbeta, dres, None, None, None, None, None


def batch_norm_nhwc(x: Tensor, gamma: Tensor, beta: Tensor, *, res: Tensor | None = None, relu: bool = False, run_mean: Tensor | None = None,
                    run_var: Tensor | None = None, eps: float = 1e-5, momentum: float = 0.1) -> Tensor:
    """[relu](batch_norm(x) [+ res]) with per-call batch statistics over all but the last axis of a contiguous tensor; the
    running statistics (if given) are updated in place like nn.BatchNorm2d in training mode."""
    if x.shape[-1] % 4: raise ValueError(f'batch_norm_nhwc: channels must be a multiple of 4, got {x.shape[-1]}')
    if res is not None and res.shape != x.shape: raise ValueError('batch_norm_nhwc: residual shape mismatch')
    return _BatchNormNHWC.apply(_f32c(x), _f32c(gamma), _f32c(beta), _f32c(res), run_mean, run_var, bool(relu), float(eps), float(momentum))


class _MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        L.require_cuda(x, what='maxpool3x3s2')
        N, H, W, Cc = x.shape
        with torch.cuda.device(x.device):
            y = torch.empty((N, (H - 1)//2 + 1, (W - 1)//2 + 1, Cc), dtype=torch.float32, device=x.device)
            idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
            L.check(L.lib().stv_maxpool3x3s2_fwd(N, H, W, Cc, L.ptr(x), L.ptr(y), L.ptr(idx), L.stream()), 'stv_maxpool3x3s2_fwd')
        ctx.save_for_backward(idx)
        ctx.shape = (N, H, W, Cc)
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        N, H, W, Cc = ctx.shape
        dy = _f32c(dy)
        with torch.cuda.device(dy.device):
            dx = torch.empty(ctx.shape, dtype=torch.float32, device=dy.device)
            L.check(L.lib().stv_maxpool3x3s2_bwd(N, H, W, Cc, L.ptr(dy), L.ptr(idx), L.ptr(dx), L.stream()), 'stv_maxpool3x3s2_bwd')
        return dx


def maxpool3x3s2(x: Tensor) -> Tensor:
    """F.max_pool2d(kernel 3, stride 2, padding 1) on a channels-last (N,H,W,C) tensor, C a multiple of 4."""
    if x.ndim != 4 or x.shape[-1] % 4: raise ValueError(f'maxpool3x3s2: expected (N,H,W,C) with C % 4 == 0, got {tuple(x.shape)}')
    return _MaxPool3x3s2.apply(_f32c(x))


class _Head3x3(torch.autograd.Function):
    """act(reflect-padded 3x3 convolution to one channel + bias): the decoder's disparity heads, as a per-pixel dot product."""
    @staticmethod
    def forward(ctx, x, w, b, act):
        L.require_cuda(x, w, b, what='head3x3')
        N, H, W, Cc = x.shape
        w_phys = w.permute(0, 2, 3, 1).contiguous()  # (1,3,3,C)
        with torch.cuda.device(x.device):
            y = torch.empty((N, H, W, 1), dtype=torch.float32, device=x.device)
            L.check(L.lib().stv_head3x3_fwd(N, H, W, Cc, L.ptr(x), L.ptr(w_phys), L.ptr(b), L.ACT[act], L.ptr(y), L.stream()), 'stv_head3x3_fwd')
        ctx.save_for_backward(x, w_phys, y)
        ctx.act, ctx.has_bias = act, b is not None
        return y

    @staticmethod
    def backward(ctx, dA):
        x, w_phys, y = ctx.saved_tensors
        N, H, W, Cc = x.shape
        dA = _f32c(dA)
        with torch.cuda.device(x.device):
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            dw = torch.zeros_like(w_phys) if ctx.needs_input_grad[1] else None
            db = torch.zeros(1, dtype=torch.float32, device=x.device) if ctx.has_bias and dw is not None else None
            dz = torch.empty((N, H, W), dtype=torch.float32, device=x.device)  # scratch: dA * act'(y), shared by both gradient kernels
            L.check(L.lib().stv_head3x3_bwd(N, H, W, Cc, L.ptr(x), L.ptr(w_phys), L.ptr(dA), L.ptr(y), L.ACT[ctx.act], L.ptr(dx), L.ptr(dw),
                                            L.ptr(db), L.ptr(dz), L.stream()), 'stv_head3x3_bwd')
        return dx, None if dw is None else dw.permute(0, 3, 1, 2), db, None


def head3x3(x: Tensor, w: Tensor, b: Tensor | None, act: str | None = 'sigmoid') -> Tensor:
    """x (N,H,W,C) channels-last, w the nn.Conv2d weight (1,C,3,3) with padding_mode='reflect' -> (N,H,W,1)."""
    Cc = x.shape[-1]
    if x.ndim != 4 or w.shape != (1, Cc, 3, 3): raise ValueError(f'head3x3: bad shapes {tuple(x.shape)}, {tuple(w.shape)}')
    if act == 'gelu': raise ValueError('head3x3: unsupported activation gelu')
    return _Head3x3.apply(_f32c(x), w, _f32c(b), act)


def head3x3_supported(C: int, out_ch: int) -> bool:
    return out_ch == 1 and 4 <= C <= 128 and (C & (C - 1)) == 0


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) on (M, K) rows: one tcgen05 GEMM forward (bias + activation in the epilogue), up to two backward."""
    @staticmethod
    def forward(ctx, x, w, b, act):
        y = gemm_tf32(x, w, bias=b, act=act)
        ctx.save_for_backward(x, w, y)
        ctx.act, ctx.has_bias = act, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        db = torch.zeros(w.shape[0], dtype=torch.float32, device=y.device) if ctx.has_bias else None
        dz = act_bwd(dy, y, ctx.act, db)
        dx = gemm_tf32(dz, w, b_mn=True) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = torch.zeros_like(w)
            gemm_tf32(dz, x, a_mn=True, b_mn=True, out=dw, accumulate=True, split_k=_split_k(w.shape[0], w.shape[1], x.shape[0]))
        return dx, dw, db, None


def linear(x: Tensor, w: Tensor, b: Tensor | None = None, act: str | None = None) -> Tensor:
    """act(x @ w.T + b): x (M,K), w (N,K), K and N multiples of 4; act in {None, relu, elu, sigmoid} (derivative from the output)."""
    if x.ndim != 2 or w.ndim != 2 or x.shape[1] != w.shape[1]: raise ValueError(f'linear: bad shapes {tuple(x.shape)}, {tuple(w.shape)}')
    if act == 'gelu': raise ValueError('linear: GELU needs the pre-activation; use convnext_mlp.')
    return _Linear.apply(_f32c(x), _f32c(w), _f32c(b), act)


def _split_k(out_rows: int, out_cols: int, k: int) -> int:
    """Reduction splits for a weight-gradient product: enough CTAs for ~2 waves of 148 SMs, >= 4 k-blocks of 32 per split."""
    tiles = ((out_rows + 127)//128)*((out_cols + 127)//128)
    return max(1, min((2*148)//tiles, k//128))  # floor: tiles*splits <= 296 CTAs = two full waves, no 1-CTA tail wave


class _ConvNeXtMlp(torch.autograd.Function):
    """out = res + gamma * (GELU(x W1^T + b1) W2^T + b2) on (M, C) rows — the pointwise half of a timm ConvNeXtBlock
    (`mlp.fc1`, GELU, `mlp.fc2`, `gamma`, residual; the reference builds it at src/networks/depth.py:97).

    Two tcgen05 GEMMs forward (bias+GELU and bias+layer-scale+residual fused into their epilogues) and four backward
    (GELU' fused into the fc2 data gradient; both weight gradients as split-K products accumulated with red.global.add)."""
    @staticmethod
    def forward(ctx, x, res, w1, b1, w2, b2, gamma, link=None):
        ctx.link = link
        M, Cc = x.shape
        z = torch.empty((M, w1.shape[0]), dtype=torch.float32, device=x.device)
        h = torch.empty_like(z)
        gemm_tf32(x, w1, bias=b1, act='gelu', aux=z, out=h)
        out = gemm_tf32(h, w2, bias=b2, gamma=gamma, res=res)
        ctx.save_for_backward(x, z, h, w1, w2, b2, gamma)
        ctx.w1_sink, ctx.b1_sink = _sink(w1), _sink(b1)
        ctx.tail_sinks = (_sink(w2), _sink(b2), _sink(gamma))
        return out

    @staticmethod
    def backward(ctx, g):
        x, z, h, w1, w2, b2, gamma = ctx.saved_tensors
        g = _f32c(g)
        M, Cc = x.shape
        Hd = w1.shape[0]
        lib = L.lib()
        with torch.cuda.device(g.device):
            w2g = torch.empty_like(w2)                                       # (C, 4C): layer-scale folded into fc2
            L.check(lib.stv_rowscale(Cc, Hd, L.ptr(w2), L.ptr(gamma), L.ptr(w2g), L.stream()), 'stv_rowscale')
        db1 = ctx.b1_sink if ctx.b1_sink is not None else torch.zeros(Hd, dtype=torch.float32, device=g.device)
        if DETERMINISTIC:   # the fused column sums take one atomic per row tile: not reproducible
            dz = gemm_tf32(g, w2g, b_mn=True, dact='gelu', dact_src=z)
            L.check(L.lib().stv_colsum(dz.shape[0], dz.shape[1], dz.shape[1], L.ptr(dz), L.ptr(db1), L.stream()), 'stv_colsum')
        else: dz = gemm_tf32(g, w2g, b_mn=True, dact='gelu', dact_src=z, colsum=db1)  # (M, 4C) = (g W2g) * GELU'(z); db1 = its column sums
        dx = gemm_tf32(dz, w1, b_mn=True) if ctx.needs_input_grad[0] else None
        w2s, b2s, gas = ctx.tail_sinks
        sunk = w2s is not None and b2s is not None and gas is not None
        # Both weight-gradient products and the layer-scale tail are off the critical path: side stream when every result is sunk.
        with (_side_stream(dz, x, g, h) if (sunk and ctx.w1_sink is not None and ctx.b1_sink is not None) else _side_stream()):
            dw1 = ctx.w1_sink if ctx.w1_sink is not None else torch.zeros_like(w1)
            gemm_tf32(dz, x, a_mn=True, b_mn=True, out=dw1, accumulate=True, split_k=_split_k(Hd, Cc, M))
            if ctx.w1_sink is not None: dw1 = None
            if ctx.b1_sink is not None: db1 = None
            G = torch.zeros_like(w2)                                             # g^T h, before the layer-scale
            gemm_tf32(g, h, a_mn=True, b_mn=True, out=G, accumulate=True, split_k=_split_k(Cc, Hd, M))
            gs = colsum(g)
            if sunk: dw2, db2, dga = w2s, b2s, gas                               # accumulate straight into the flat gradient buffer
            else: dw2, db2, dga = torch.zeros_like(w2), torch.zeros_like(b2), torch.zeros_like(gamma)
            with torch.cuda.device(g.device):
                L.check(lib.stv_ls_tail(Cc, Hd, L.ptr(G), L.ptr(w2), L.ptr(b2), L.ptr(gamma), L.ptr(gs), L.ptr(dw2), L.ptr(db2), L.ptr(dga),
                                        L.stream()), 'stv_ls_tail')
        g_res = g if ctx.needs_input_grad[1] else None
        if g_res is not None and ctx.link is not None:  # handed to the block's depthwise data-gradient kernel (see _DwConv7.backward)
            ctx.link['g_res'] = g_res
            g_res = None
        if sunk: return (dx, g_res, dw1, db1, None, None, None, None)
        return (dx, g_res, dw1, db1, dw2, db2, dga, None)


def convnext_mlp(x: Tensor, res: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, gamma: Tensor, link: dict | None = None) -> Tensor:
    """x, res: (M, C) contiguous rows (channels-last pixels); w1 (4C, C), w2 (C, 4C). -> (M, C)."""
    if x.ndim != 2 or res.shape != x.shape or w1.shape[1] != x.shape[1] or w2.shape != w1.shape[::-1]:
        raise ValueError(f'convnext_mlp: bad shapes {tuple(x.shape)}, {tuple(res.shape)}, {tuple(w1.shape)}, {tuple(w2.shape)}')
    return _ConvNeXtMlp.apply(_f32c(x), _f32c(res), _f32c(w1), _f32c(b1), _f32c(w2), _f32c(b2), _f32c(gamma), link)


class _SmoothLossEx(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img, use_edges: bool, use_laplacian: bool, use_blur: bool):
        L.require_cuda(disp, img, what='smooth_loss_ex')
        b, _, H, W = disp.shape
        Cc = img.shape[1]
        lib, dev = L.lib(), disp.device
        flags = (int(use_edges), int(use_laplacian), int(use_blur))
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            dg, ig = torch.empty_like(disp), torch.empty_like(disp)
            ws = _ws(lib.stv_smooth_ex_workspace_bytes(b, Cc, H, W), dev)
            L.check(lib.stv_smooth_ex_fwd(b, Cc, H, W, *flags, L.ptr(disp), L.ptr(img), L.ptr(loss), L.ptr(dg), L.ptr(ig), L.ptr(ws),
                                          ws.numel(), L.stream()), 'stv_smooth_ex_fwd')
        ctx.flags, ctx.shape = flags, (b, Cc, H, W)
        ctx.save_for_backward(disp, ws)
        ctx.mark_non_differentiable(dg, ig)
        return loss, dg, ig

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2):
        disp, ws = ctx.saved_tensors
        with torch.cuda.device(disp.device):
            g = torch.empty_like(disp)
            L.check(L.lib().stv_smooth_ex_bwd(*ctx.shape, *ctx.flags, L.ptr(disp), L.ptr(g_loss.to(torch.float32).contiguous()), L.ptr(g),
                                              L.ptr(ws), ws.numel(), L.stream()), 'stv_smooth_ex_bwd')
        return g, None, None, None, None


def smooth_loss_ex(disp: Tensor, img: Tensor, *, use_edges: bool = False, use_laplacian: bool = False, use_blur: bool = False):
    """`SmoothReg.forward` with every constructor flag (src/regularizers/smooth.py:51-97), single scale: disp (b,1,H,W), img (b,C,H,W)
    -> (loss, disp_grad (b,1,H,W), image_grad (b,1,H,W)); differentiable in `disp`."""
    if disp.ndim != 4 or disp.shape[1] != 1 or img.ndim != 4 or img.shape[0] != disp.shape[0] or img.shape[-2:] != disp.shape[-2:]:
        raise ValueError(f'Non-matching shapes. ({tuple(disp.shape)} vs. {tuple(img.shape)})')
    return _SmoothLossEx.apply(_f32c(disp), _f32c(img.detach()), bool(use_edges), bool(use_laplacian), bool(use_blur))


class _FeatReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, img, order: int, use_edges: bool):
        L.require_cuda(feat, img, what='feat_reg')
        b, Cc, H, W = feat.shape
        Ci = img.shape[1]
        lib, dev = L.lib(), feat.device
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            fg = torch.empty_like(feat)
            ws = _ws(lib.stv_feat_reg_workspace_bytes(b, Cc, Ci, H, W), dev)
            L.check(lib.stv_feat_reg_fwd(b, Cc, Ci, H, W, order, int(use_edges), L.ptr(feat), L.ptr(img), L.ptr(loss), L.ptr(fg), L.ptr(ws),
                                         ws.numel(), L.stream()), 'stv_feat_reg_fwd')
        ctx.args = (b, Cc, Ci, H, W, order, int(use_edges))
        ctx.save_for_backward(feat, ws)
        ctx.mark_non_differentiable(fg)
        return loss, fg

    @staticmethod
    def backward(ctx, g_loss, _g):
        feat, ws = ctx.saved_tensors
        with torch.cuda.device(feat.device):
            g = torch.empty_like(feat)
            L.check(L.lib().stv_feat_reg_bwd(*ctx.args, L.ptr(feat), L.ptr(g_loss.to(torch.float32).contiguous()), L.ptr(g), L.ptr(ws),
                                             ws.numel(), L.stream()), 'stv_feat_reg_bwd')
        return g, None, None, None


def feat_reg(feat: Tensor, img: Tensor, *, order: int, use_edges: bool = False):
    """FeatPeakReg (order 1) / FeatSmoothReg (order 2) forward (src/regularizers/smooth.py:100-176): feat (b,C,H,W), img (b,Ci,H,W)
    -> (loss, feat_grad (b,C,H,W)); differentiable in `feat`."""
    if feat.ndim != 4 or img.ndim != 4 or feat.shape[0] != img.shape[0] or feat.shape[-2:] != img.shape[-2:]:
        raise ValueError(f'Non-matching shapes. ({tuple(feat.shape)} vs. {tuple(img.shape)})')
    return _FeatReg.apply(_f32c(feat), _f32c(img.detach()), int(order), bool(use_edges))


class _PointwiseReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kind: int, sign: float):
        L.require_cuda(x, what='pointwise_reg')
        lib, dev = L.lib(), x.device
        with torch.cuda.device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ws = _ws(lib.stv_regr_workspace_bytes(), dev)
            L.check(lib.stv_pwreg_fwd(x.numel(), kind, float(sign), L.ptr(x), L.ptr(loss), L.ptr(ws), ws.numel(), L.stream()), 'stv_pwreg_fwd')
        ctx.args = (x.numel(), kind, float(sign))
        ctx.save_for_backward(x)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        x, = ctx.saved_tensors
        with torch.cuda.device(x.device):
            g = torch.empty_like(x)
            L.check(L.lib().stv_pwreg_bwd(*ctx.args, L.ptr(x), L.ptr(g_loss.to(torch.float32).contiguous()), L.ptr(g), L.stream()), 'stv_pwreg_bwd')
        return g, None, None


def mean_reg(x: Tensor, sign: float = 1.0) -> Tensor:
    """OccReg (src/regularizers/occlusion.py:9-40): sign * mean(x)."""
    return _PointwiseReg.apply(_f32c(x), 0, float(sign))


def bce_to_one(x: Tensor) -> Tensor:
    """MaskReg (src/regularizers/mask.py:11-30): F.binary_cross_entropy(x, ones_like(x))."""
    return _PointwiseReg.apply(_f32c(x), 1, 1.0)


# ---------------------------------------------------------------------------------------------------------------------
# Bilinear resampling (aspect-ratio augmentation)
# ---------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def resample_bilinear(x: Tensor, size: tuple[int, int], *, mode: str, ax: float = 0., bx: float = 0., ay: float = 0., by: float = 0.) -> Tensor:
    """(..., H, W) fp32 -> (..., oh, ow) on the libstv kernel (include/stv.h: stv_resample_bilinear). No autograd (augmentation).

    mode 'interp': F.interpolate(size, mode='bilinear', align_corners=False).
    mode 'grid':   sample positions ix = ax*j + bx, iy = ay*i + by in pixel units, zero padding (affine_grid + grid_sample)."""
    L.require_cuda(x, what='resample_bilinear')
    if x.ndim < 2: raise ValueError(f'resample_bilinear: expected (..., H, W), got {tuple(x.shape)}')
    if mode not in ('interp', 'grid'): raise ValueError(f'resample_bilinear: unknown mode {mode!r}')
    x = _f32c(x)
    H, W = x.shape[-2:]
    oh, ow = int(size[0]), int(size[1])
    if oh <= 0 or ow <= 0: raise ValueError(f'resample_bilinear: bad output size {size}')
    P = x.numel()//(H*W)
    with torch.cuda.device(x.device):
        out = torch.empty((*x.shape[:-2], oh, ow), dtype=torch.float32, device=x.device)
        if mode == 'interp': ax, ay, bx, by = W/ow, H/oh, 0., 0.
        for p0 in range(0, P, 65535):  # gridDim.z limit
            pn = min(65535, P - p0)
            L.check(L.lib().stv_resample_bilinear(pn, H, W, oh, ow, ax, bx, ay, by, 1 if mode == 'interp' else 0,
                                                  x.data_ptr() + p0*H*W*4, out.data_ptr() + p0*oh*ow*4, L.stream()), 'stv_resample_bilinear')
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Logging statistics
# ---------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def mean_std(tensors: list[Tensor]) -> Tensor:
    """(k,2) device tensor {mean, unbiased std} of k fp32 device tensors in one launch pair (include/stv.h: stv_mean_std). No sync:
    read it back with ONE `.cpu()` when logging (the reference's summarize_depth syncs once per statistic, trainer.py:486-503)."""
    if not tensors: raise ValueError('mean_std: no tensors')
    L.require_cuda(*tensors, what='mean_std')
    ts = [_f32c(t) for t in tensors]
    dev = ts[0].device
    with torch.cuda.device(dev):
        out = torch.empty((len(ts), 2), dtype=torch.float32, device=dev)
        for k0 in range(0, len(ts), 16):
            part = ts[k0:k0 + 16]
            ws = _ws(L.lib().stv_mean_std_workspace_bytes(len(part)), dev)
            counts = (C.c_longlong*len(part))(*[t.numel() for t in part])
            L.check(L.lib().stv_mean_std(len(part), L.ptr_array(part), counts, out.data_ptr() + k0*8, L.ptr(ws), ws.numel(), L.stream()),
                    'stv_mean_std')
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Optimiser step
# ---------------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def adamw_step_(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, *, n_decay: int, lr: float, beta1: float,
                beta2: float, eps: float, weight_decay: float, step: int, grad_scale: float = 1.0) -> None:
    """In-place fused AdamW on flat fp32 buffers; elements [0, n_decay) get weight decay."""
    L.require_cuda(param, grad, exp_avg, exp_avg_sq, what='adamw_step_')
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not t.is_contiguous() or t.numel() != param.numel(): raise ValueError('adamw_step_: buffers must be flat, contiguous, same size.')
    with torch.cuda.device(param.device):
        L.check(L.lib().stv_adamw_step(L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), param.numel(), n_decay,
                                       lr, beta1, beta2, eps, weight_decay, grad_scale, step, L.stream()), 'stv_adamw_step')
