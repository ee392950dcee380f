"""Running turtleFSI problem files UNMODIFIED ("The problem-file API ... stay unchanged, so the new path is
a drop-in").

A turtleFSI problem file is written against DOLFIN: `from dolfin import *`, `from turtleFSI.problems import *`,
`from turtleFSI.modules import *`, hooks that receive DOLFIN objects (`DVP.sub(1)`, `boundaries`, `dvp_["n"].sub(0,
deepcopy=True)`, `ds(6)`, `n`). `install()` puts stand-ins for those three packages (this directory) on `sys.path`
— only when no real DOLFIN is importable — and `load_problem_file(path)` returns the file wrapped in `DolfinDialect`,
which hands every hook the DOLFIN-flavoured view of the driver's namespace and turns what the hooks return back into
what turtlefsi_b200.monolithic.run works with. The hot path is untouched by this: matrices and residuals are
assembled by the sm_100a kernels behind the C-ABI."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def real_dolfin_available():
    try:
        spec = importlib.util.find_spec("dolfin")
    except (ImportError, ValueError):
        spec = None
    return spec is not None and not os.path.abspath(spec.origin or "").startswith(_HERE)


def install(force=False):
    """make `import dolfin`, `import mpi4py`, `import turtleFSI` resolve to the stand-ins of this directory."""
    if _HERE in sys.path:
        return True
    if not force and real_dolfin_available():
        return False
    sys.path.insert(0, _HERE)
    for name in ("dolfin", "mpi4py", "turtleFSI"):
        m = sys.modules.get(name)
        if m is not None and not os.path.abspath(getattr(m, "__file__", "") or "").startswith(_HERE):
            raise RuntimeError("%s is already imported from %s" % (name, getattr(m, "__file__", "?")))
    return True


def load_problem_file(path):
    """import a problem file by path; DOLFIN-dialect files (they import `dolfin`) come back wrapped."""
    path = os.path.abspath(path)
    src = open(path).read()
    uses_dolfin = ("from dolfin" in src) or ("import dolfin" in src)
    if uses_dolfin and not install():
        raise RuntimeError("%s is written against DOLFIN and a real DOLFIN is importable: run it with the reference "
                           "and the binding of INTEGRATION.md" % path)
    name = "_tfsi_problem_" + os.path.splitext(os.path.basename(path))[0]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    sys.path.insert(0, os.path.dirname(path))              # problem files import their neighbours (problem_aneu_2solid.py)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(os.path.dirname(path))
    return DolfinDialect(mod) if uses_dolfin else mod


class DolfinDialect:
    """the seven hooks of a DOLFIN-dialect problem file, callable with turtlefsi_b200's namespace."""

    def __init__(self, mod):
        self.mod = mod
        self.cmesh = self.cdomains = self.cbnd = None
        self._DVP = None
        self._fn = {}

    def __getattr__(self, name):                 # any other attribute of the file (classes, helpers)
        return getattr(self.mod, name)

    # ---- hooks --------------------------------------------------------------------------------------------
    def set_problem_parameters(self, **ns):
        return self.mod.set_problem_parameters(**ns)

    def get_mesh_domain_and_boundaries(self, **ns):
        import numpy as np
        import dolfin
        mesh, domains, boundaries = self.mod.get_mesh_domain_and_boundaries(**ns)
        if boundaries.facets is None or domains.facets is not None:
            raise TypeError("get_mesh_domain_and_boundaries must return (mesh, cell MeshFunction, facet MeshFunction)")
        self.cmesh, self.cdomains, self.cbnd = mesh, domains, boundaries
        native = mesh.native
        native.markers = np.asarray(domains.array(), dtype=np.int32)
        dolfin._CTX.update(mesh=mesh, boundaries=boundaries, domains=domains)
        return native, native.markers, boundaries.facets

    def _view(self, ns):
        """the DOLFIN-flavoured namespace: what `vars()` of turtleFSI/monolithic.py holds when it calls a hook."""
        import dolfin
        v = dict(ns)
        v.update(mesh=self.cmesh, domains=self.cdomains, boundaries=self.cbnd)
        if "DVP" in ns:
            if self._DVP is None or self._DVP.native is not ns["DVP"]:
                self._DVP = dolfin.FunctionSpace(self.cmesh, native=ns["DVP"])
            DVP = self._DVP
            v["DVP"] = DVP
            dolfin._CTX.update(space=ns["DVP"])
            tests = dolfin.TestFunctions(DVP)
            v.update(zip(("phi", "psi", "gamma", "beta"), tests))
        if "assembler" in ns:
            dolfin._CTX.update(assembler=ns["assembler"])
        if "dvp_" in ns and "DVP" in ns:
            asm = ns["assembler"]
            for key, vec in ns["dvp_"].items():
                f = self._fn.get(key)
                if f is None or f.vec is not vec:
                    self._fn[key] = dolfin.Function(v["DVP"], vec=vec, assembler=asm)
            v["dvp_"] = dict(self._fn)
            names = ("d_", "v_", "p_", "w_")
            for i, nm in enumerate(names[:v["DVP"].num_sub_spaces()]):
                v[nm] = {key: f.sub(i) for key, f in self._fn.items()}
        v["n"] = dolfin.FacetNormal(self.cmesh)                                             # monolithic.py:82
        v["k"] = dolfin.Constant(ns["dt"])
        v["ds"] = dolfin.Measure("ds", domain=self.cmesh, subdomain_data=self.cbnd)        # :120-122
        v["dS"] = dolfin.Measure("dS", domain=self.cmesh, subdomain_data=self.cbnd)
        v["dx"] = dolfin.Measure("dx", domain=self.cmesh, subdomain_data=self.cdomains)
        if "dx_f_id_list" in ns:                                                            # modules/domain.py:54-103
            v["dx_f"] = [v["dx"](i) for i in ns["dx_f_id_list"]]
            v["dx_s"] = [v["dx"](i) for i in ns["dx_s_id_list"]]
        return v

    def _call(self, name, ns):
        out = getattr(self.mod, name)(**self._view(ns))
        if out:
            # DOLFIN-flavoured objects the hook merely passed through must not overwrite the driver's own
            for key in ("mesh", "domains", "boundaries", "DVP", "dvp_", "d_", "v_", "p_", "w_"):
                out.pop(key, None) if isinstance(out, dict) else None
        return out

    def initiate(self, **ns):
        return self._call("initiate", ns)

    def create_bcs(self, **ns):
        return self._call("create_bcs", ns)

    def pre_solve(self, **ns):
        return self._call("pre_solve", ns)

    def post_solve(self, **ns):
        return self._call("post_solve", ns)

    def finished(self, **ns):
        if ns.get("results_folder") is None:     # the reference always has one (create_folders); library callers may not
            return None
        return self._call("finished", ns)
