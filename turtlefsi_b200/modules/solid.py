"""turtleFSI/modules/solid.py counterpart: selects the solid terms assembled by the CUDA kernels
(solid.py:41-43, 62-68 St.Venant-Kirchhoff; zero forms of no_solid.py:14-15)."""


def solid_setup(**namespace):
    term = dict(solid="solid")
    return dict(F_solid_linear=term, F_solid_nonlinear=term)
