"""`from turtleFSI.modules import *` (turtleFSI/modules/__init__.py: `from .common import *`), and the physics
modules under their reference names (turtleFSI.modules.fluid ... newtonsolver)."""
import importlib as _il
import sys as _sys

from turtlefsi_b200.modules.common import *  # noqa: F401,F403

for _n in ("common", "fluid", "no_fluid", "solid", "no_solid", "laplace", "biharmonic", "elastic", "no_extrapolation",
           "newtonsolver"):
    _sys.modules[__name__ + "." + _n] = _il.import_module("turtlefsi_b200.modules." + _n)
