"""Import shim: the package directory is named ``arnoldimethod.jl_b200`` (with a dot, after
the reference's name), which Python cannot import by name - load it by path and expose it as
``b200arnoldi``:

    import b200arnoldi as b2a
    decomp, history = b2a.partialschur(A, nev=10, tol=1e-6, which="SR")
"""

import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "arnoldimethod.jl_b200")
_NAME = "arnoldimethod_jl_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
    )
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
package = _pkg
