from .argpar import *  # noqa: F401,F403
