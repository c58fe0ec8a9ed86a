"""SASS evidence per kernel of libtreedet.so (cuobjdump -sass): instruction counts of the mnemonics that prove what a
kernel does on sm_100a -- TMA (UTMALDG / UTMASTG), mbarrier (SYNCS.*), warp reductions (REDUX), popcount (POPC),
funnel shifts (SHF), shared / global accesses, float64 arithmetic, votes / shuffles, local-memory traffic (spills).

    python profiles/sass_summary.py [path/to/libtreedet.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "treedetection_b200", "csrc", "libtreedet.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
GROUPS = [("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("REDUX", r"\bC?REDUX"),
          ("POPC", r"\bPOPC"), ("SHF", r"\bSHF"), ("VOTE", r"\bVOTE"), ("SHFL", r"\bSHFL"), ("MATCH", r"\bMATCH"),
          ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("ATOMS", r"\bATOMS"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"),
          ("LD.generic", r"\bLD\b|\bLD\.E"), ("ST.generic", r"\bST\b|\bST\.E"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"),
          ("DFMA/DADD/DMUL", r"\bD(FMA|ADD|MUL)\b"), ("FFMA/FADD/FMUL", r"\bF(FMA|ADD|MUL)\b"), ("HMMA/UTCMMA", r"MMA"),
          ("BAR", r"\bBAR\b"), ("WARPSYNC", r"\bWARPSYNC")]
kern, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if not m:
        continue
    ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
    kern[cur]["total"] += 1
    for name, pat in GROUPS:
        if re.match(pat, ins):
            kern[cur][name] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(kern)} sm_100a kernels (cuobjdump -sass); static instruction counts per kernel")
print("# tensor-core instructions (HMMA / UTCMMA) are absent by design: nothing on this path is a contraction\n")
for name, c in sorted(kern.items(), key=lambda kv: -kv[1]["total"]):
    short = re.sub(r"\(anonymous namespace\)::", "", demangle(name))
    short = re.sub(r"\((int|bool|unsigned int|StatsMode)\)", "", short)
    short = re.sub(r"\(.*", "", short)
    if "cub" in short or "thrust" in short:
        continue
    cols = "  ".join(f"{k} {c[k]}" for k, _ in GROUPS if c[k])
    print(f"{short[:70]:70s} total {c['total']:6d}  {cols}")
