"""Importable alias for the `s-volsdf_b200/` package directory (a hyphen is not a valid module name).

`import svolsdf_b200` executes `s-volsdf_b200/__init__.py` with this module's `__path__` pointing at
that directory, so `svolsdf_b200.model.network.VolSDFNetwork` resolves to the real sources.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 's-volsdf_b200')
__path__ = [_real]
__file__ = _os.path.join(_real, '__init__.py')
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, 'exec'))
