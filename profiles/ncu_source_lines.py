"""Dynamic instruction counts of one kernel per SOURCE line: joins the SASS page of an .ncu-rep captured with
`--set full --import-source on` (executed instructions per SASS address) with the line table of the same build
(`nvdisasm --print-line-info` on the cubin inside libtreedet.so; the library is built with -lineinfo).

    python profiles/ncu_source_lines.py gpurun_out/x.ncu-rep simplify_kernel geometry > profiles/rNN_source_x.txt
                                        report               kernel regex    cubin (source file stem)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "treedetection_b200", "csrc", "libtreedet.so")
CSRC = os.path.join(ROOT, "treedetection_b200", "csrc")


def line_table(cubin_stem, kernel_rx):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", cubin_stem, LIB], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=d, check=True, capture_output=True,
                             text=True).stdout.splitlines()
    start = end = None
    for i, l in enumerate(dis):
        if l.strip().startswith(".text."):
            if start is None and re.search(kernel_rx, l):
                start = i
            elif start is not None:
                end = i
                break
    table, cur = {}, None
    for l in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main(rep, kernel_rx, cubin_stem):
    table = line_table(cubin_stem, kernel_rx)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_rx],
                         capture_output=True, text=True).stdout
    tables = []
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == "Kernel Name":
            tables.append({"name": r[1], "rows": []})
        elif r and r[0] == "Address":
            tables[-1]["hdr"] = r
        elif r and tables:
            tables[-1]["rows"].append(r)
    src = {}
    for t in tables:
        h = {k: i for i, k in enumerate(t["hdr"])}
        inst, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
        base = None
        for r in t["rows"]:
            a = int(r[h["Address"]], 16) if r[h["Address"]].startswith("0x") else int(r[h["Address"]])
            base = a if base is None else base
            key = table.get(a - base)
            inst[key] += int(r[h["Instructions Executed"]] or 0)
            thr[key] += int(r[h["Thread Instructions Executed"]] or 0)
            smp[key] += int(r[h["# Samples"]] or 0)
        tot, ts = sum(inst.values()), max(1, sum(smp.values()))
        print(f"== {t['name'][:110]}")
        print(f"== {tot} warp instructions executed; share, live lanes and stall samples per source line (>= 0.4 %)")
        for key, v in sorted(inst.items(), key=lambda kv: -kv[1]):
            if v < 0.004 * tot:
                break
            text = ""
            if key:
                if key[0] not in src:
                    p = os.path.join(CSRC, key[0])
                    src[key[0]] = open(p).read().splitlines() if os.path.exists(p) else []
                text = src[key[0]][key[1] - 1].strip()[:90] if key[1] <= len(src[key[0]]) else ""
            name = f"{key[0]}:{key[1]}" if key else "?"
            print(f"{100 * v / tot:5.1f}%  lanes {thr[key] / max(v, 1):4.1f}  samples {100 * smp[key] / ts:4.1f}%  {name:<24} {text}")
        print()


if __name__ == "__main__":
    main(*sys.argv[1:4])
