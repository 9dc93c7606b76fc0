"""Importable alias of the product package.

The product lives in ``lyft-3d-object-detection_b200/`` (the name the build
contract fixes); a hyphenated directory cannot be imported directly, so this
shim points its ``__path__`` there and executes that package's ``__init__``.
``import lyft3d_b200.bev`` therefore loads
``lyft-3d-object-detection_b200/bev.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "lyft-3d-object-detection_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
