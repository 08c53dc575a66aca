"""Import shim: the package sources live in ``cross-modal-video-engine_b200/`` (a directory name
Python cannot import directly).  This module extends its search path to that directory and runs
the real ``__init__`` there, so ``import cross_modal_video_engine_b200.evaluation`` etc. work."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "cross-modal-video-engine_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
