"""Importable alias for the package directory ``delayed-streams-modeling_b200/``.

The directory name the project layout asks for contains hyphens, which Python
cannot import directly; this stub points ``dsm_b200``'s search path at it so
``import dsm_b200.mimi`` resolves to ``delayed-streams-modeling_b200/mimi.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "delayed-streams-modeling_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
