import os as _os
import sys as _sys

__path__ = [_os.path.dirname(__file__)]
_rel = __name__.replace(".", _os.sep)
for _p in _sys.path:
    _cand = _os.path.join(_p, _rel)
    if _os.path.isdir(_cand) and _os.path.abspath(_cand) != _os.path.abspath(__path__[0]):
        __path__.append(_cand)
