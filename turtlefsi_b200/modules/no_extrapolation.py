"""turtleFSI/modules/no_extrapolation.py counterpart: no lifting term is added to F_fluid_linear."""


def extrapolate_setup(F_fluid_linear, **namespace):
    return dict(F_fluid_linear=F_fluid_linear)
