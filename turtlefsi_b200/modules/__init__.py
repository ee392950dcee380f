"""Drop-in counterparts of turtleFSI/modules: the same module names and function names
(`fluid_setup`, `solid_setup`, `extrapolate_setup`, `solver_setup`, `newtonsolver`) selected by
the same `--fluid/--solid/--extrapolation/--solver` values (turtleFSI/monolithic.py:139-154), but
returning TERM DESCRIPTORS for the B200 assembler instead of UFL forms."""
