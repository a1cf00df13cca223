"""Import shim: the product package lives in `real-time-recommendation-system-with-feature-store_b200/` (a directory
name Python cannot import directly); `import b200rec` resolves its submodules from there."""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "real-time-recommendation-system-with-feature-store_b200")
__path__ = [_PKG]
with open(_os.path.join(_PKG, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_PKG, "__init__.py"), "exec"))
