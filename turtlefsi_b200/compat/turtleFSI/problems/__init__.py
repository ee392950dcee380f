"""`from turtleFSI.problems import *` (turtleFSI/problems/TF_fsi.py:19): default_variables and the default hooks."""
from turtlefsi_b200.problems import *  # noqa: F401,F403
from turtlefsi_b200.problems import _compiler_parameters, default_variables  # noqa: F401


def info_blue(s, check=True):
    print(s)


info_green = info_red = info_blue
