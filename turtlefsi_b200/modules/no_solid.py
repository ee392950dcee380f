"""turtleFSI/modules/no_solid.py counterpart: zero forms over the solid regions (no_solid.py:14-15)."""
from ..ufl_lite import Form


def solid_setup(**namespace):
    return dict(F_solid_linear=Form(), F_solid_nonlinear=Form())
