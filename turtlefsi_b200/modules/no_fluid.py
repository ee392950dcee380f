"""turtleFSI/modules/no_fluid.py counterpart: zero forms over the fluid regions (no_fluid.py:14-15) — no kernel term,
so nothing is assembled there and `ident_zeros` puts unit rows, as in the reference."""
from ..ufl_lite import Form


def fluid_setup(**namespace):
    return dict(F_fluid_linear=Form(), F_fluid_nonlinear=Form())
