"""Import shim for the REAL reference (jspenmar/slowtv_monodepth): the source tree at /root/reference in the build container,
or its byte-compiled build `oracle/_ref/` (oracle/build_ref.py) on the GPU box, where the source tree does not exist.

TEST INFRASTRUCTURE ONLY. Used by `oracle/make_golden.py`, by the tests that execute the reference's own PyTorch code (CPU
cross-checks here; `tests/test_plugin_gpu.py` on the GPU box) and by `bench.py --impl reference`. Nothing in the product
package imports it.

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

SRC_ROOT = Path(os.environ.get('STV_REFERENCE_ROOT', '/root/reference'))
BUILT_ROOT = Path(__file__).resolve().parent/'_ref'
BUILT_ARCHIVE = BUILT_ROOT/'ref_build.zip'   # src/**.pyc, imported through zipimport (oracle/build_ref.py)


def _root() -> Path | None:
    if (SRC_ROOT/'src'/'tools'/'geometry.py').is_file(): return SRC_ROOT
    if BUILT_ARCHIVE.is_file(): return BUILT_ARCHIVE
    return None


REF_ROOT = _root() or SRC_ROOT


def available() -> bool:
    return _root() is not None


def kind() -> str:
    """'source' (the reference tree itself), 'bytecode' (oracle/_ref built from it) or 'absent'."""
    r = _root()
    return 'absent' if r is None else ('source' if r == SRC_ROOT else 'bytecode')


def load_cfg(*names: str) -> dict:
    """The reference's experiment configurations, merged in order like `src.utils.io.load_merge_yaml` (api/train/train.py:29)."""
    import json
    load()
    from src.utils import io
    if kind() == 'source':
        return io.load_merge_yaml(*[SRC_ROOT/'cfg'/n for n in names])
    cfgs = json.loads((BUILT_ROOT/'cfg.json').read_text())
    out: dict = {}
    for n in names: out = _merge(out, cfgs[n])
    return out


def _merge(old: dict, new: dict) -> dict:
    """Recursive dict merge, new values win (the rule of src/utils/io.py:134-161)."""
    out = dict(old)
    for k, v in new.items():
        out[k] = _merge(out[k], v) if (isinstance(v, dict) and isinstance(out.get(k), dict)) else v
    return out


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
    if not available(): raise FileNotFoundError(f'Reference not found: neither {SRC_ROOT} nor the built {BUILT_ROOT}')
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
