"""Importable alias for the `multiagent-quadruped-environment_b200/` package directory.

The product directory carries the reference's repository name (with hyphens, so it cannot be
imported directly); this shim points `mqe_b200.*` at it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "multiagent-quadruped-environment_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
