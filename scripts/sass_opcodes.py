"""SASS opcode summary of the shipped library (cuobjdump -sass): per hot kernel the instruction count, the FP64 / memory /
atomic / TMA opcodes and the forms of the scatter instructions.  python scripts/sass_opcodes.py > profiles/rN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "turtlefsi_b200/libturtlefsi_b200.so"
WANT = ["residual_kernel<3, false, 16, 0>", "jac_fluid_kernel<3, 14>", "jac_solid_kernel<3, 14, 0, 0>", "jac_solid_kernel<3, 14, 0, 1>",
        "bitmap_stream_kernel<1>", "lin_jacobian_kernel<3>", "ident_zeros_kernel", "sumsq_kernel"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
print("SASS opcode summary of %s (cuobjdump -sass, sm_100a)\n" % LIB)
for name, body in zip(names, blocks):
    if not any(w in name for w in WANT):
        continue
    ops = collections.Counter()
    forms = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body):
        full = m.group(1)
        ops[full.split(".")[0]] += 1
        if full.startswith(("REDG", "ATOMG", "UBLKCP", "RED.")):
            forms[full] += 1
    n = sum(ops.values())
    print(name)
    print("  %d instructions" % n)
    print("  DFMA %d, DMUL %d, DADD %d, REDG %d, ATOMG %d, ATOMS %d, UBLKCP (TMA bulk copy) %d, SYNCS (mbarrier) %d, LDS %d, STS %d, "
          "LDG %d, STG %d, LDC %d, LDCU %d, LDL %d, STL %d, BAR %d, HMMA/UTCMMA (tensor) %d"
          % (tuple(ops[k] for k in ("DFMA", "DMUL", "DADD", "REDG", "ATOMG", "ATOMS", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG",
                                    "LDC", "LDCU", "LDL", "STL", "BAR")) + (ops["HMMA"] + ops["UTCMMA"],)))
    if forms:
        print("  " + ", ".join("%s x%d" % kv for kv in sorted(forms.items())))
    print()
