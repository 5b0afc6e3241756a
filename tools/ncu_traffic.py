"""profiles/kernel_traffic.json from an `ncu --set full` report: DRAM bytes (read + written) of one launch of every
kernel captured, keyed by the names bench.py looks up.
usage: python tools/ncu_traffic.py report.ncu-rep config frames layout > profiles/kernel_traffic.json"""
import csv
import json
import subprocess
import sys


def main(rep, config, frames, layout):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def gb(r, k):
        v = float(r[idx[k]].replace(",", ""))
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[idx[k]]]
    d, detail = {}, {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        name = "fnp::" + name if not name.startswith("fnp::") else name
        rd, wr = gb(r, "dram__bytes_read.sum"), gb(r, "dram__bytes_write.sum")
        if name not in d:
            d[name] = rd + wr
            detail[name] = "%.1f MB read + %.1f MB written" % (rd / 1e6, wr / 1e6)
    print(json.dumps({"config": config, "frames": int(frames), "layout": layout, "dram_bytes_per_launch": d,
                      "detail": detail,
                      "source": "%s: ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of ONE "
                                "launch of each kernel inside `bench.py --frames %s --layout %s` (resident arm)" % (rep, frames, layout)},
                     indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
