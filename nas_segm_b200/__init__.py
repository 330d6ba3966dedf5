"""Importable alias of the product package, whose directory name (``nas-segm-pytorch_b200/``) is not a valid
Python identifier.  ``import nas_segm_b200`` executes that package's ``__init__`` with its directory as the
package path, so ``nas_segm_b200.nn``, ``.engine``, ``.helpers``, ``.rl`` resolve into it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nas-segm-pytorch_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
