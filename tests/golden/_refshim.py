"""Import shim for the *unmodified* reference package (only usable where /root/reference exists).

Used ONLY by tests/golden/make_golden.py (fixture generation in the build container).
Nothing in the product, the `-m gpu` tests, smoke() or bench.py imports this file.

The reference imports five modules that are not installed (dotmap, omegaconf, wget, matplotlib,
requests.packages.target — xnrs/training.py:3, xnrs/utils.py:144); they are stubbed, the
reference itself is untouched.
"""
import sys
import types

REF_ROOT = '/root/reference'


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


def load_reference():
    def stub(name, **attrs):
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
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import xnrs  # noqa: F401
    return xnrs
