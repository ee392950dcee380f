"""Entry point — turtleFSI/run_turtle.py + the head of turtleFSI/monolithic.py:20-43: parse the command line, find the
problem file (a local file first, then the shipped problems), drop the arguments that were not given, and run the
monolithic solver with the CUDA assembler behind the reference's plugin interface.

    python -m turtlefsi_b200 --problem TF_fsi --extrapolation laplace --theta 0.51 -dt 0.01 -T 0.05
    python -m turtlefsi_b200 -p /path/to/turtleFSI/problems/TF_fsi.py -T 0.05        (an UNMODIFIED reference file)
"""
import os
import sys


def make_b200(linear_solver):
    from .assembler import FSIAssembler

    def make_assembler(mesh, space, regions, ns):
        return FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                            extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                            quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"], gravity=ns["gravity"])

    def make_solver(assembler):
        from .linear_solver import HostLU, make_device_solver
        if linear_solver in ("superlu", "host"):
            return HostLU(assembler.rowptr.cpu().numpy(), assembler.colidx.cpu().numpy(), assembler.num_dofs)
        return make_device_solver(assembler)[0]
    return make_assembler, make_solver


def make_oracle():
    """the CPU checker behind the same interface — for tests of the host logic on machines without a GPU; never the
    product path (`--backend oracle` must be asked for explicitly)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle.backend import OracleAssembler
    from .linear_solver import HostLU

    def make_assembler(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        ext, sub = ns["extrapolation"], ns["extrapolation_sub_type"]
        c_port = ext in ("no_extrapolation", "biharmonic") or (ext == "laplace" and sub == "constant")   # else numpy (slow)
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"],
                               10 if mesh.gdim == 2 else 6, use_c_port=c_port)
    return make_assembler, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs)


def main(argv=None):
    from . import monolithic
    from .utils.argpar import parse
    args = parse(argv)
    overrides = {k: v for k, v in args.__dict__.items() if v is not None}          # monolithic.py:38-40
    problem = overrides.pop("problem")
    backend = overrides.pop("backend")
    if backend == "b200":
        make_assembler, make_solver = make_b200(overrides.get("linear_solver", "mumps"))
    else:
        make_assembler, make_solver = make_oracle()
    ns = monolithic.run(problem, make_assembler, make_solver, overrides=overrides, create_results_folder=True, progress=True)
    return ns


if __name__ == "__main__":
    main()
