"""`get_class`: the reference's plug-in mechanism (volsdf/utils/general.py:10-16) — the model is named by a
dotted string in the config (`train.model_class`), imported and instantiated with `conf=`."""
import importlib


def get_class(kls):
    module, _, name = kls.rpartition('.')
    return getattr(importlib.import_module(module), name)
