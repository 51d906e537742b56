"""Import shim: ``import velocyto_b200`` loads the package that lives in ``velocyto.py_b200/``.

The package directory keeps the name the project layout prescribes (``velocyto.py_b200``),
which is not a valid Python identifier; this module turns itself into that package
(``__path__`` points at the directory, ``__init__.py`` runs in this namespace) so that
``velocyto_b200.estimation`` etc. import normally.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "velocyto.py_b200")
__path__ = [_pkg_dir]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
