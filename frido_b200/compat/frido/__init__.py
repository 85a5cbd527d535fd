# shim package: B200-native hot-path modules first, the reference checkout (if on sys.path) for the rest
import os as _os
import sys as _sys

__path__ = [_os.path.dirname(__file__)]
for _p in _sys.path:
    _cand = _os.path.join(_p, __name__)
    if _os.path.isdir(_cand) and _os.path.abspath(_cand) != _os.path.abspath(__path__[0]):
        __path__.append(_cand)
