"""Importable alias of the `pytv-4d_b200/` package directory (a hyphen cannot appear in a module name).

`import pytv_b200 as pytv` gives the reference's GPU-path module layout: `pytv.tv_GPU`, `pytv.tv_operators_GPU`.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pytv-4d_b200")
__path__.insert(0, _impl)
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _f
