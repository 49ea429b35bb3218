"""Import alias for the package directory `three-mlagents_b200/` (a hyphen is not importable).

`import three_mlagents_b200` executes `three-mlagents_b200/__init__.py` with this module's
`__path__` pointing at that directory, so `three_mlagents_b200.training`, `.registry`, `.cli`,
`.vec_env`, `.native` ... all resolve to the files there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "three-mlagents_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
