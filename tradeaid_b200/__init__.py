"""Import alias: the product package lives in ``trade-aid-multimodal-transformer_b200/`` (the layout the project
brief names); a hyphen cannot appear in a Python package name, so this shim exposes it as ``tradeaid_b200``."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "trade-aid-multimodal-transformer_b200")
__path__.insert(0, _REAL)
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
