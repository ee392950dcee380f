"""turtleFSI/modules/no_extrapolation.py counterpart: the mesh-lifting term added to F_fluid_linear
(laplace.py:33-46 / biharmonic.py:31-36 cell terms / no_extrapolation.py)."""


def extrapolate_setup(F_fluid_linear, extrapolation_sub_type="constant", **namespace):
    term = dict(F_fluid_linear)
    term.update(extrapolation="no_extrapolation", extrapolation_sub_type=extrapolation_sub_type)
    return dict(F_fluid_linear=term)
