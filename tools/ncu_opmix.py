"""Opcode mix of one kernel from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_opmix.py report.ncu-rep kernel-substring [top]"""
import collections, csv, subprocess, sys, io

def main(rep, kname, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[]); blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if kname not in b["name"]:
            continue
        hdr = b["rows"][0]
        iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
        iSmp = hdr.index("# Samples")
        op, tot, recs = collections.Counter(), 0, []
        for r in b["rows"][1:]:
            try:
                n = int(r[iE])
            except Exception:
                continue
            s = r[iS].strip(); t = s.split()
            o = t[1] if t[0].startswith("@") else t[0]
            op[o.split(".")[0]] += n; tot += n
            recs.append((n, int(r[iSmp] or 0), s))
        print("==", b["name"][:90], "warp-instr:", tot)
        for k, v in op.most_common(top):
            print("  %-10s %12d %5.1f%%" % (k, v, 100.0 * v / tot))
        print("  -- top stall-sample lines")
        for n, smp, s in sorted(recs, key=lambda x: -x[1])[:12]:
            print("  %8d smp %10d exec  %s" % (smp, n, s[:90]))
        return

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
