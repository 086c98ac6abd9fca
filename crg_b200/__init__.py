"""Import shim: the package directory is ``conservativeregridding.jl_b200`` (not a valid
Python identifier), so ``crg_b200`` extends its ``__path__`` there and re-exports it."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "conservativeregridding.jl_b200")
__path__.insert(0, _PKG)

with open(_os.path.join(_PKG, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG, "__init__.py"), "exec"))
