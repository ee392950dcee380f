from turtlefsi_b200.utils.argpar import *  # noqa: F401,F403
