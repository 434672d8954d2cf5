"""Drop-in installation into the reference's registry (src/registry.py) — the plugin boundary of SURVEY 8b.

    import slowtv_monodepth_b200.plugin as plugin
    plugin.install()                      # before MonoDepthModule(cfg) is built
    runpy.run_path('api/train/train.py')  # the reference's entry script, byte-identical

`install()` (a) triggers the reference's lazy registrations so the originals exist, (b) re-registers the B200 classes under
the same keys with `overwrite=True` (registry.py:131-132), and (c) rebinds the names the trainer resolves at call time:
`src.core.handlers.image_recon / feat_recon / disp_smooth` (trainer.py:389,402,437), `src.core.trainer.ViewSynth` (trainer.py:168) and
`src.core.trainer.aspect_ratio_aug` (trainer.py:12,54-60; the GPU augmentation of SURVEY 8f).
With `fast_step=True` (default) it also removes the host synchronisations of the reference's own `MonoDepthModule.step`, which
otherwise cap the drop-in's speed whatever the kernels do: `forward` (trainer.py:192-278; ATen's host-synchronising
`T.inverse()`, device-tensor comparisons on `supp_idxs`) and `forward_postprocess` (trainer.py:280-348; four ATen upsamples +
`to_scaled`) are rebound to the sync-free `MonoDepthStep` versions (one libstv kernel per scale), `summarize_depth / _pose / _K`
(trainer.py:486-529, ~28 `.item()` syncs) to one multi-tensor statistics kernel returning device scalars, and the module timer
is built with `sync_gpu=False` (trainer.py:69 syncs 18x per step). `step`, `forward_loss`, `training_step` stay the reference's
own code, and the step becomes capturable in a CUDA graph (`graphed_step`).
The reference tree itself is not modified. Requires the reference to be importable (`src` on sys.path).
"""
from __future__ import annotations

import functools

from . import aspect_ratio, geometry, handlers, losses, networks, regularizers

__all__ = ['install', 'uninstall', 'graphed_step', 'REPLACED']

REPLACED = {
    'net': {'depth': networks.DepthNet, 'pose': networks.PoseNet},
    'dec': {'monodepth': networks.MonodepthDecoder},
    # the reference registers ONE class under three keys (src/losses/reconstruction.py:12)
    'loss': {'img_recon': losses.ReconstructionLoss, 'feat_recon': losses.ReconstructionLoss, 'autoenc_recon': losses.ReconstructionLoss,
             'depth_regr': losses.RegressionLoss, 'stereo_const': losses.RegressionLoss,   # src/losses/regression.py:40
             'disp_smooth': regularizers.SmoothReg, 'feat_peaky': regularizers.FeatPeakReg, 'feat_smooth': regularizers.FeatSmoothReg,
             'disp_mask': regularizers.MaskReg, 'disp_occ': regularizers.OccReg},
}
_saved: dict = {}


def _fast_step_patches(ref_trainer) -> dict:
    """Sync-free replacements for methods of the reference's MonoDepthModule (same names, same return contracts)."""
    from . import trainer as T_
    ref_post = ref_trainer.MonoDepthModule.forward_postprocess

    def forward(self, x):
        if set(self.nets) - {'depth', 'pose'}: return _saved[('method', 'forward')](self, x)  # autoencoder etc.: reference code
        return T_.MonoDepthStep.forward(self, x)

    def forward_postprocess(self, fwd, x, y):
        plain = all(not any(t in k for t in ('mask', 'stereo', 'autoenc')) for k in fwd if isinstance(k, str))
        if not plain or '_idxs' not in fwd: return ref_post(self, fwd, x, y)
        return T_.MonoDepthStep.forward_postprocess(self, fwd, x, y)

    def _summary(self, fwd, keep):
        s = T_.summarize(fwd)
        return {n: s.values[i] for i, n in enumerate(s.names) if keep(n)}   # 0-dim device tensors: no host sync until logged

    return {
        'forward': forward,
        'forward_postprocess': forward_postprocess,
        'summarize_depth': lambda self, fwd: _summary(self, fwd, lambda n: n.startswith(('disp_', 'depth_'))),
        'summarize_pose': lambda self, fwd: _summary(self, fwd, lambda n: n.startswith('T_')),
        'summarize_K': lambda self, fwd: _summary(self, fwd, lambda n: n in ('fx', 'fy', 'cx', 'cy')),
    }


def install(nets: bool = True, loss: bool = True, fast_step: bool = True) -> None:
    import src.registry as reg
    reg.trigger_nets(); reg.trigger_decoders(); reg.trigger_losses()
    import src.core.handlers as ref_handlers
    import src.core.trainer as ref_trainer

    def put(kind: str, key: str, cls) -> None:
        _saved.setdefault(('reg', kind, key), reg._REG[kind].get(key))
        reg.register(key, type=kind, overwrite=True)(cls)

    if nets:
        for kind in ('net', 'dec'):
            for key, cls in REPLACED[kind].items(): put(kind, key, cls)
    if loss:
        for key, cls in REPLACED['loss'].items(): put('loss', key, cls)
        _saved.setdefault(('attr', 'image_recon'), ref_handlers.image_recon)
        _saved.setdefault(('attr', 'disp_smooth'), ref_handlers.disp_smooth)
        _saved.setdefault(('attr', 'feat_recon'), ref_handlers.feat_recon)
        _saved.setdefault(('attr', 'stereo_const'), ref_handlers.stereo_const)
        _saved.setdefault(('attr', 'depth_regr'), ref_handlers.depth_regr)
        for name in ('feat_smooth', 'disp_occ', 'disp_mask'): _saved.setdefault(('attr', name), getattr(ref_handlers, name))
        _saved.setdefault(('attr', 'ViewSynth'), ref_trainer.ViewSynth)
        ref_handlers.image_recon = handlers.image_recon
        ref_handlers.disp_smooth = handlers.disp_smooth
        ref_handlers.feat_recon = handlers.feat_recon
        ref_handlers.stereo_const = handlers.stereo_const
        ref_handlers.depth_regr = handlers.depth_regr
        ref_handlers.feat_smooth, ref_handlers.disp_occ, ref_handlers.disp_mask = handlers.feat_smooth, handlers.disp_occ, handlers.disp_mask
        ref_trainer.ViewSynth = geometry.ViewSynth
    # MonoDepthModule.__init__ binds `aspect_ratio_aug` by name from src.core.trainer (trainer.py:12,54-60)
    _saved.setdefault(('attr', 'aspect_ratio_aug'), ref_trainer.aspect_ratio_aug)
    ref_trainer.aspect_ratio_aug = aspect_ratio.aspect_ratio_aug
    if fast_step and nets and loss:
        for name, fn in _fast_step_patches(ref_trainer).items():
            _saved.setdefault(('method', name), getattr(ref_trainer.MonoDepthModule, name))
            setattr(ref_trainer.MonoDepthModule, name, fn)
        timer_cls = _saved.setdefault(('attr', 'MultiLevelTimer'), ref_trainer.MultiLevelTimer)

        @functools.wraps(timer_cls)
        def timer_no_sync(*a, **k):
            k['sync_gpu'] = False
            return timer_cls(*a, **k)
        ref_trainer.MultiLevelTimer = timer_no_sync


def graphed_step(module, opt, example_batch, warmup: int = 3):
    """The reference's own `MonoDepthModule.step` (built through its registry after `install()`) + backward, captured as ONE CUDA
    graph and replayed per batch: `runner.run(batch)` = load inputs, replay, all-reduce, AdamW. `module` only needs `.step`."""
    from .trainer import GraphedTrainStep
    return GraphedTrainStep(module, opt, example_batch, warmup=warmup)


def uninstall() -> None:
    import src.registry as reg
    import src.core.handlers as ref_handlers
    import src.core.trainer as ref_trainer
    for key, val in list(_saved.items()):
        if key[0] == 'reg':
            if val is None: reg._REG[key[1]].pop(key[2], None)
            else: reg._REG[key[1]][key[2]] = val
        elif key[0] == 'method': setattr(ref_trainer.MonoDepthModule, key[1], val)
        elif key[1] in ('image_recon', 'disp_smooth', 'feat_recon', 'stereo_const', 'depth_regr', 'feat_smooth', 'disp_occ', 'disp_mask'): setattr(ref_handlers, key[1], val)
        else: setattr(ref_trainer, key[1], val)
    _saved.clear()
