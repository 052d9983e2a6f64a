"""Attribute-access dict used for configs and loaded models.  Stand-in for `easydict.EasyDict`
(the reference's config files do `from easydict import EasyDict as edict`; reference
config/infer_config.py:1).  `install_easydict_shim()` registers it under that module name when
the real package is absent, so unmodified model-folder configs keep loading."""
import sys
import types


class AttrDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        super().__setitem__(k, v)

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def install_easydict_shim():
    try:
        import easydict  # noqa: F401
    except ImportError:
        m = types.ModuleType('easydict')
        m.EasyDict = AttrDict
        sys.modules['easydict'] = m
