"""Import shim for the REAL reference tree (jspenmar/slowtv_monodepth at /root/reference).

TEST INFRASTRUCTURE ONLY. This module is used by `oracle/make_golden.py` (and by the optional
`-m "not gpu"` cross-checks that skip when the reference tree is absent) to execute the reference's own
PyTorch code on CPU inside the build container, so that the oracle restatement in `oracle/` can be pinned
and golden vectors can be generated. It never travels to the GPU box (the reference tree does not exist
there) and nothing in the product package imports it.

The reference has import-time dependencies that are not installed here (matplotlib, skimage, kornia, timm,
torchmetrics, lmdb, pytorch_lightning). They are only needed for plotting / logging / data loading, none of
which is on the hot path, so we insert placeholder modules before `import src`.  The timm placeholder
delegates `create_model` to the oracle's own restatement of the timm feature extractors
(`oracle/nets.py`), i.e. the reference's DepthNet/PoseNet/MonodepthDecoder code runs *unchanged* on top of it.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types
from pathlib import Path

REF_ROOT = Path(os.environ.get('STV_REFERENCE_ROOT', '/root/reference'))


def available() -> bool:
    return (REF_ROOT/'src'/'tools'/'geometry.py').is_file()


def _try(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # Behave as a package so that `import a.b` works.
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules: setattr(sys.modules[parent], child, m)
    return m


def _raiser(what: str):
    def fn(*a, **k): raise RuntimeError(f'"{what}" is a placeholder: dependency not installed in this container.')
    return fn


def install_stubs() -> None:
    """Insert placeholder modules for the reference's missing import-time dependencies."""
    import torch.nn as nn

    if not _try('matplotlib.pyplot'):
        _mod('matplotlib'); _mod('matplotlib.pyplot', Axes=object, Figure=object)
        _mod('matplotlib.cm'); _mod('matplotlib.colors')
    if not _try('skimage.feature'):
        _mod('skimage'); _mod('skimage.feature', canny=_raiser('skimage.feature.canny'))
    if not _try('kornia.filters'):
        _mod('kornia')
        _mod('kornia.filters', gaussian_blur2d=_raiser('kornia.filters.gaussian_blur2d'))
        _mod('kornia.geometry'); _mod('kornia.geometry.transform', center_crop=_raiser('kornia center_crop'))
        _mod('kornia.augmentation', ColorJiggle=object, ColorJitter=object, RandomHorizontalFlip=object)
    if not _try('lmdb'):
        _mod('lmdb')
    if not _try('torchmetrics'):
        class Metric(nn.Module):
            def __init__(self, *a, **k): super().__init__()
            def add_state(self, name, default, dist_reduce_fx=None): self.register_buffer(name, default)
        _mod('torchmetrics', Metric=Metric)
    if not _try('pytorch_lightning'):
        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k): pass
            def log_dict(self, *a, **k): pass
            def log(self, *a, **k): pass
        class _Empty:
            def __init__(self, *a, **k): pass
        _mod('pytorch_lightning', LightningModule=LightningModule, Trainer=_Empty, Callback=_Empty,
             seed_everything=lambda *a, **k: None)
        _mod('pytorch_lightning.callbacks', Callback=_Empty, TQDMProgressBar=_Empty, RichProgressBar=_Empty,
             ModelCheckpoint=_Empty, LearningRateMonitor=_Empty, EarlyStopping=_Empty,
             StochasticWeightAveraging=_Empty)
        _mod('pytorch_lightning.loggers', WandbLogger=_Empty, TensorBoardLogger=_Empty)
        _mod('pytorch_lightning.utilities'); _mod('pytorch_lightning.utilities.rank_zero', rank_zero_only=lambda f: f)
    if not _try('timm'):
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
        from oracle import nets as onets
        _mod('timm', create_model=onets.timm_create_model)
        _mod('timm.optim'); _mod('timm.optim.optim_factory', create_optimizer_v2=onets.timm_create_optimizer_v2)


_LOADED = False


def load():
    """Make `import src` resolve to the reference tree. Returns the imported `src` package."""
    global _LOADED
    if not available(): raise FileNotFoundError(f'Reference tree not found at {REF_ROOT}')
    sys.dont_write_bytecode = True  # The reference tree is read-only.
    if not _LOADED:
        install_stubs()
        if str(REF_ROOT) not in sys.path: sys.path.insert(0, str(REF_ROOT))
        lvl = logging.root.manager.disable
        logging.disable(logging.WARNING)  # `src/paths.py` warns about a missing PATHS.yaml.
        try:
            import src  # noqa
        finally:
            logging.disable(lvl)
        _LOADED = True
    import src
    return src
