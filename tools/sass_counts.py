"""SASS mnemonic counts per kernel of libfnp_sm100.so (cuobjdump -sass): the packed fp32x2 instructions
(FFMA2 / FMUL2 / FADD2), TMA bulk copies (UBLKCP), warp reductions (REDUX), shared / global atomics.
usage: python tools/sass_counts.py [lib.so] > profiles/rNN_sass_counts.txt"""
import collections
import re
import subprocess
import sys

OPS = ["FFMA2", "FMUL2", "FADD2", "UBLKCP", "SYNCS", "REDUX", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "MUFU", "F2I",
       "BAR", "FFMA", "FMUL", "FADD", "FMNMX", "FMNMX3"]


def main(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "")
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    print("%-44s %7s " % ("kernel (sm_100a SASS, static counts)", "instr") + " ".join("%6s" % o for o in OPS))
    for k, c in per.items():
        print("%-44s %7d " % (k[:44], c["_total"]) + " ".join("%6d" % c[o] for o in OPS))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "findnpropagate_b200/libfnp_sm100.so")
