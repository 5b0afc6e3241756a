"""Key metrics per kernel launch of an .ncu-rep (raw page): time, DRAM bytes, issue, occupancy."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__inst_executed.sum", "Minst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%")]


def conv(v, unit, tag):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if tag in ("rdMB", "wrMB"):
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
    if tag == "us":
        x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(unit, 1.0)
    if tag == "Minst":
        x *= 1e-6
    return "%.1f" % x


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-44s " % "kernel" + " ".join("%8s" % t for _, t in KEYS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:44]
        print("%-44s " % name + " ".join("%8s" % (conv(r[idx[k]], units[idx[k]], t) if k in idx else "-") for k, t in KEYS))


if __name__ == "__main__":
    main(sys.argv[1])
