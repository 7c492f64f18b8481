"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference package from `oracle/_ref/xnrs` (see oracle/make_ref.py).

Used by `bench.py --impl reference` and bench.py's `cpu_baseline` legs only.  The reference imports five modules that are
not installed in this image (dotmap, omegaconf, wget, matplotlib, requests.packages.target — xnrs/training.py:3,
xnrs/utils.py:144); they are stubbed here, the reference files themselves are untouched.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


class DotMap(dict):
    """attribute-access dict; a missing key reads as an empty (falsy) DotMap like the real package."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return self[k] if k in self else DotMap()

    __setattr__ = dict.__setitem__


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'xnrs', '__init__.py'))


_loaded = None


def load():
    """-> (xnrs package, xnrs.training module, make_model, metrics module) of the reference copy; raises if it is absent"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('oracle/_ref/xnrs is missing: run `python oracle/make_ref.py` where /root/reference exists')

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    stub('dotmap', DotMap=DotMap)
    stub('omegaconf', DictConfig=dict)
    stub('wget')
    mpl = stub('matplotlib')
    mpl.pyplot = stub('matplotlib.pyplot')
    import requests.packages as rp
    rp.target = None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import xnrs
    import xnrs.training as T
    from xnrs.evaluation import metrics
    from xnrs.models import make_model
    _loaded = (xnrs, T, make_model, metrics)
    return _loaded


def make_trainer(cfg: dict, trainer: str = 'ContrastiveRankingTrainer'):
    """a reference trainer around a reference model WITHOUT its DataLoaders (BaseTrainer.__init__, training.py:26-44, minus
    `_init_dataloaders`): the bench feeds `_train_step` / `_test_step` pre-built batches"""
    import torch
    _, T, make_model, _ = load()
    cfg = DotMap(cfg)
    model = make_model(cfg)
    tr = object.__new__(getattr(T, trainer))
    tr.cfg, tr.model = cfg, model
    tr.device = torch.device(cfg.device)
    tr.model.to(tr.device)
    tr.optimizer = torch.optim.Adam(tr.model.parameters(), lr=cfg.lr)          # training.py:39
    tr._init_loss()
    tr.current_epoch = tr.current_train_step = tr.current_test_step = 0
    return tr
