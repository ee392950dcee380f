"""Command line of the drop-in — turtleFSI/utils/argpar.py:121-323 re-hosted on argparse (configargparse is not in the
image; `-c/--config` reads the same `key = value` lines). Same switches, same short forms, same defaults (None = "keep
the problem file's value"), same treatment of unknown `--key=value` arguments (:295-314: literal_eval, repeated keys
become lists) and of `--new-arguments` / `--compiler-parameters` key=value pairs."""
import argparse
import ast

import yaml

__all__ = ["parse", "str2bool", "restricted_float", "StoreDictKeyPair"]


class StoreDictKeyPair(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        my_dict = {}
        for kv in values:
            k, v = kv.split("=")
            try:
                my_dict[k] = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                my_dict[k] = v
        setattr(namespace, self.dest, my_dict)


def str2bool(boolean):
    if boolean.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if boolean.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected, not {}.'.format(boolean))


def restricted_float(x):
    x = float(x)
    if x < 0.0 or x > 1.0:
        raise argparse.ArgumentTypeError("{} not in range [0.0, 1.0]".format(x))
    return x


def _config_args(path):
    out = []
    for line in open(path):
        line = line.split("#")[0].strip()
        if not line:
            continue
        k, _, v = line.partition("=")
        out.append("--%s=%s" % (k.strip().lstrip("-"), v.strip()) if v else "--" + k.strip().lstrip("-"))
    return out


def parse(argv=None):
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = argparse.ArgumentParser(
        prog="python -m turtlefsi_b200",
        description="turtleFSI's monolithic FSI solver with the Newton hot path (residual / Jacobian assembly) on B200: "
                    "same problem files, same switches.")
    A = parser.add_argument
    A('-c', '--config', default=None, help='configuration file (key = value lines)')
    A("-p", "--problem", type=str, default="turtle_demo", metavar="Problem",
      help="problem file: a name shipped with the package (TF_cfd, TF_csm, TF_fsi, turtle_demo, ...) or a local file, "
           "which may be an unmodified turtleFSI problem file written against DOLFIN")
    A("-t", "--theta", type=restricted_float, default=None, metavar="Theta")
    A("-f", "--fluid", type=str, default=None, metavar="Fluid", help="'fluid', 'no_fluid' or the name of your own module")
    A("-s", "--solid", type=str, default=None, metavar="Solid", help="'solid', 'no_solid' or the name of your own module")
    A("-e", "--extrapolation", type=str, default=None, metavar="Extrapolation method",
      help="'laplace', 'elastic', 'biharmonic', 'no_extrapolation' or the name of your own module")
    A("-et", "--extrapolation-sub-type", type=str, metavar="Extrapolation sub type", default=None,
      choices=["constant", "small_constant", "volume", "volume_change", "constrained_disp", "constrained_disp_vel"])
    A("--bc-ids", nargs="+", type=int, default=None, metavar="ID list")
    for name in ("Um", "rho-f", "mu-f", "rho-s", "mu-s", "nu-s", "lambda-s", "gravity"):
        A("--" + name, type=float, default=None)
    A("--solid_properties", type=yaml.safe_load, default=None, help="Solid properties as a dictionary")
    A("--fluid_properties", type=yaml.safe_load, default=None, help="Fluid properties as a dictionary")
    A("--dx-f-id", type=int, default=None)
    A("--dx-s-id", type=int, default=None)
    A("--linear-solver", type=str, default=None, help="'mumps' (the reference's name) selects the default of this "
      "package: device LU; 'superlu' the host LU")
    A("--atol", type=float, default=None)
    A("--rtol", type=float, default=None)
    A("--max-it", type=int, default=None)
    A("--lmbda", type=restricted_float, default=None)
    A("--recompute", type=int, default=None)
    A("--recompute-tstep", type=int, default=None)
    A("--compiler-parameters", dest="compiler_parameters", action=StoreDictKeyPair, nargs="+", metavar="KEY=VAL", default=None)
    A("-v", "--verbose", type=str2bool, default=None)
    A("--loglevel", type=int, default=None)
    A("--save-step", type=int, default=None)
    A("--save-deg", type=int, default=None)
    A("--checkpoint-step", type=int, default=None)
    A("--folder", type=str, default=None)
    A("--sub-folder", type=str, default=None)
    A("--restart-folder", type=str, default=None)
    A("-dt", "--time-step", metavar="Time step", type=float, default=None)
    A("-T", "--end-time", type=float, metavar="End time", default=None)
    A("--p-deg", type=int, default=None)
    A("--v-deg", type=int, default=None)
    A("--d-deg", type=int, default=None)
    A("--killtime", type=int, default=None)
    A("--new-arguments", action=StoreDictKeyPair, nargs="+", metavar="KEY=VAL", default=None)
    # not in the reference: which assembler / linear solver stand behind the hooks
    A("--backend", choices=["b200", "oracle"], default="b200",
      help="b200: the CUDA path (default, needs a GPU, no fallback); oracle: the CPU checker under oracle/ (tests only)")

    if "-c" in argv or "--config" in argv or any(a.startswith("--config=") for a in argv):
        pre, _ = parser.parse_known_args(argv)
        argv = _config_args(pre.config) + argv
    args, unknownargs = parser.parse_known_args(argv)
    d = args.__dict__
    d.pop("config", None)
    if d["solid_properties"]:
        for k, v in d["solid_properties"].items():
            if k != "material_model" and k != "dx_s_id":
                d["solid_properties"][k] = float(v)
    if d["fluid_properties"]:
        for k, v in d["fluid_properties"].items():
            d["fluid_properties"][k] = float(v)
    d.update({'dt': d.pop('time_step'), 'T': d.pop('end_time')})
    if args.new_arguments is not None:
        d.update(args.new_arguments)
    d.pop("new_arguments", None)
    unknownargs_dict = {}
    for arg in unknownargs:
        k, v = arg.lstrip("-").split('=')
        try:
            val = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            val = v
        if k in unknownargs_dict:
            if not isinstance(unknownargs_dict[k], list):
                unknownargs_dict[k] = [unknownargs_dict[k]]
            unknownargs_dict[k] = unknownargs_dict[k] + [val]
        else:
            unknownargs_dict[k] = [val] if isinstance(val, list) else val
    d.update(unknownargs_dict)
    if args.compiler_parameters is not None:
        from turtlefsi_b200.problems import default_variables
        cp = dict(default_variables["compiler_parameters"])
        cp.update(args.compiler_parameters)
        args.compiler_parameters = cp
    return args
