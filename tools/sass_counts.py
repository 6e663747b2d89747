"""Static SASS opcode counts per kernel of the shipped library (profiles/r2_sass_tcgen05.txt).
Usage: python tools/sass_counts.py [path/to/libcleanrl_cuda.so] > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "MUFU", "LDS", "STS",
       "BAR", "ATOM", "RED", "LDL", "STL"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "cleanrl.jl_b200", "libcleanrl_cuda.so")
    txt = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
    print("# SASS instruction counts of the shipped libcleanrl_cuda.so (sm_100a), per kernel")
    print("# command: python tools/sass_counts.py (cuobjdump -sass, opcodes counted per function: static counts, not executed)")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,")
    print("# FFMA2/FMUL2/FADD2 = packed FP32, LDL/STL = local-memory (spill) loads/stores\n")
    print("%-72s %7s " % ("kernel", "total") + " ".join("%7s" % o for o in OPS))
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        try:
            name = subprocess.check_output(["c++filt", name], text=True).strip()
        except Exception:
            pass
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        c = collections.Counter()
        n = 0
        for l in f.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                n += 1
                c[m.group(1)] += 1
        print("%-72s %7d " % (name[:72], n) + " ".join("%7d" % c[o] for o in OPS))


if __name__ == "__main__":
    main()
