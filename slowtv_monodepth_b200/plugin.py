"""Drop-in installation into the reference's registry (src/registry.py) — the plugin boundary of SURVEY 8b.

    import slowtv_monodepth_b200.plugin as plugin
    plugin.install()                      # before MonoDepthModule(cfg) is built
    runpy.run_path('api/train/train.py')  # the reference's entry script, byte-identical

`install()` (a) triggers the reference's lazy registrations so the originals exist, (b) re-registers the B200 classes under
the same keys with `overwrite=True` (registry.py:131-132), and (c) rebinds the names the trainer resolves at call time:
`src.core.handlers.image_recon / disp_smooth` (trainer.py:389,437), `src.core.trainer.ViewSynth` (trainer.py:168) and
`src.core.trainer.aspect_ratio_aug` (trainer.py:12,54-60; the GPU augmentation of SURVEY 8f).
The reference tree itself is not modified. Requires the reference to be importable (`src` on sys.path).
"""
from __future__ import annotations

from . import aspect_ratio, geometry, handlers, losses, networks, regularizers

__all__ = ['install', 'uninstall', 'REPLACED']

REPLACED = {
    'net': {'depth': networks.DepthNet, 'pose': networks.PoseNet},
    'dec': {'monodepth': networks.MonodepthDecoder},
    'loss': {'img_recon': losses.ReconstructionLoss, 'disp_smooth': regularizers.SmoothReg},
}
_saved: dict = {}


def install(nets: bool = True, loss: bool = True) -> None:
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
        _saved.setdefault(('attr', 'ViewSynth'), ref_trainer.ViewSynth)
        ref_handlers.image_recon = handlers.image_recon
        ref_handlers.disp_smooth = handlers.disp_smooth
        ref_trainer.ViewSynth = geometry.ViewSynth
    # MonoDepthModule.__init__ binds `aspect_ratio_aug` by name from src.core.trainer (trainer.py:12,54-60)
    _saved.setdefault(('attr', 'aspect_ratio_aug'), ref_trainer.aspect_ratio_aug)
    ref_trainer.aspect_ratio_aug = aspect_ratio.aspect_ratio_aug


def uninstall() -> None:
    import src.registry as reg
    import src.core.handlers as ref_handlers
    import src.core.trainer as ref_trainer
    for key, val in list(_saved.items()):
        if key[0] == 'reg':
            if val is None: reg._REG[key[1]].pop(key[2], None)
            else: reg._REG[key[1]][key[2]] = val
        elif key[1] in ('image_recon', 'disp_smooth'): setattr(ref_handlers, key[1], val)
        else: setattr(ref_trainer, key[1], val)
    _saved.clear()
