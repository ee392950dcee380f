"""turtlefsi_b200 — B200-native hot path of turtleFSI's monolithic Newton assembly."""
__version__ = "0.1.0"
