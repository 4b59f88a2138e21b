"""Importable alias for the product package.

The product lives in ``adversarial-attacks-on-gan-based-image-fusion_b200/`` (the name the
build contract fixes); hyphens are not importable, so this shim points ``sfattack``'s
``__path__`` at that directory and executes its ``__init__``.  ``import sfattack.engine``
therefore resolves to ``adversarial-attacks-on-gan-based-image-fusion_b200/engine.py``.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
PKG_DIR = _os.path.join(_os.path.dirname(_here), "adversarial-attacks-on-gan-based-image-fusion_b200")
__path__ = [PKG_DIR]
_init = _os.path.join(PKG_DIR, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
