"""Import alias for the product package.

The package directory is ``deeprank-gnn-2_b200/`` (the name the project layout fixes),
which is not a valid Python identifier.  This stub makes it importable as
``deeprank2_b200`` by pointing ``__path__`` at that directory and executing its
``__init__.py`` in this module's namespace; sub-modules
(``deeprank2_b200.neuralnets.gnn.ginet`` ...) then resolve through ``__path__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deeprank-gnn-2_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _fh
