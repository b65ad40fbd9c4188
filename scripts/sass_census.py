"""Opcode census of the hot kernels in libmodest_b200.so (cuobjdump -sass): which memory / sync instructions the
SASS actually contains.  python scripts/sass_census.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "modest_b200", "libmodest_b200.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, ops = None, collections.defaultdict(collections.Counter)
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
dem = dict(zip(ops, names))
want = ("pp_count_kernel", "pp_entropy", "knn_select_fast", "mutual_edges", "dbscan_union", "box_beta32", "pp_join_kernel", "pp_hist_tile",
        "transform_gather", "ransac_score", "seed_nms", "box_prereject", "bev_pairs", "ground_mask_compact")
keys = ["LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "LDS.128", "STS.128", "ATOMS", "ATOMG", "REDG", "MATCH", "SHFL", "VOTE", "BAR",
        "DFMA", "MUFU", "UBLKCP", "UTMALDG", "LDGSTS", "HMMA"]
print("# SASS opcode census of the hot kernels (`cuobjdump -sass modest_b200/libmodest_b200.so`, sm_100a)\n")
print("Static instruction counts per kernel; a prefix column counts every variant (`LDG.E.128` includes `.CONSTANT` etc.).")
print("No bulk-copy / TMA (`UBLKCP`, `UTMALDG`), no `cp.async` (`LDGSTS`) and no tensor-core instructions: the path is gather- and")
print("issue-bound SIMT code (DESIGN.md 4.1); vector loads are `LDG.E.128` of float4 records.\n")
print("| kernel | total | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for f, c in sorted(ops.items(), key=lambda x: dem[x[0]]):
    name = dem[f].split("(")[0].replace("modest::", "").replace("void ", "")
    if not any(t in name for t in want):
        continue
    row = [str(sum(v for op, v in c.items() if op == k or op.startswith(k + "."))) for k in keys]
    print(f"| `{name[:58]}` | {sum(c.values())} | " + " | ".join(row) + " |")
