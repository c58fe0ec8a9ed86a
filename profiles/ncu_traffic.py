"""DRAM traffic per P1 launch from an `ncu --set full` report -> profiles/p1_traffic.json, which
bench.py quotes as roofline.traffic.  One P1 launch = the aligned-rows kernel plus the (small)
unaligned-rows kernel of the same td_tile_cut_normalize call.
usage: python profiles/ncu_traffic.py gpurun_out/x.ncu-rep profiles/rNN_ncu_x.txt"""
import csv
import json
import os
import subprocess
import sys


def main(path, summary_name):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: hdr.index(n) for n in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                     "gpu__time_duration.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if "tile_resize" not in name:
            continue
        key = "unaligned" if "<0," in name or "<false" in name else "aligned"
        rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
        per_kernel.setdefault(key, []).append((rd, wr))
    res = {"kernels": {}, "source": summary_name, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    total = 0.0
    for key, v in per_kernel.items():
        rd = sum(x[0] for x in v) / len(v)
        wr = sum(x[1] for x in v) / len(v)
        res["kernels"][key] = {"read": rd, "write": wr, "launches_captured": len(v)}
        total += rd + wr
    res["traffic_bytes_per_launch"] = total
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump(res, open(os.path.join(here, "p1_traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1]))
