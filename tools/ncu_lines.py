"""Per-CUDA-source-line totals of one kernel from an .ncu-rep captured with --import-source on:
warp instructions executed, stall samples, average active threads.
usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv
import io
import subprocess
import sys


def main(rep, kname, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kname], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, recs, cur = None, None, [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            iE, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        elif hdr is not None and r[0] != "" and r[0] not in ("Function Name",):
            try:
                cur = dict(file=cur_file, line=int(r[0]), src=r[1].strip(), inst=int(r[iE] or 0), smp=int(r[iS] or 0), thr=int(r[iT] or 0))
                recs.append(cur)
            except ValueError:
                pass
    tot = sum(x["inst"] for x in recs) or 1
    tsmp = sum(x["smp"] for x in recs) or 1
    print("kernel %s: %.1f M warp-instr, %d samples" % (kname, tot / 1e6, tsmp))
    print("%6s %6s %5s  %s" % ("inst%", "smp%", "thr", "line"))
    for x in sorted(recs, key=lambda x: -x["inst"])[:top]:
        print("%6.2f %6.2f %5.1f  %s:%d  %s" % (100.0 * x["inst"] / tot, 100.0 * x["smp"] / tsmp, x["thr"] / max(x["inst"], 1),
                                               x["file"], x["line"], x["src"][:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
