"""turtleFSI/modules/no_fluid.py counterpart: selects the fluid terms assembled by the CUDA kernels
(N1..N6, L1, L2, L4 of fluid.py:32-56; zero forms of no_fluid.py:14-15)."""


def fluid_setup(**namespace):
    term = dict(fluid="no_fluid")
    return dict(F_fluid_linear=term, F_fluid_nonlinear=term)
