"""`mpi4py` stand-in: the problem files import it only to reduce a point evaluation over ranks
(turtleFSI/problems/TF_fsi.py:17,208-219). One process per GPU here, and hooks see global vectors, so the
reduction is the identity (see dolfin.MPI.comm_world.Allreduce in this directory)."""


class _Op:
    def __init__(self, name):
        self.name = name


class MPI:
    MIN, MAX, SUM = _Op("MIN"), _Op("MAX"), _Op("SUM")
