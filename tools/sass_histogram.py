"""Per-kernel SASS opcode histogram of the built library: which kernels carry tcgen05 (UTCHMMA / UTCQMMA ...), tensor-memory access
(LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), tcgen05.commit barriers (UTCBAR) and legacy mma.sync (HMMA).

    python tools/sass_histogram.py [tubelet-transformer_b200/libtuber_b200.so] > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tubelet-transformer_b200", "libtuber_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "FFMA2", "FFMA", "MUFU", "SYNCS", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
per, cur, i = collections.OrderedDict(), None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[i]) if i < len(names) else m.group(1)
        i += 1
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in KEYS:
            per[cur][op] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
print("# " + " ".join(f"{k:>8s}" for k in ["total"] + KEYS) + "  kernel")
tot = collections.Counter()
for name, c in per.items():
    tot.update(c)
    print("  " + " ".join(f"{c.get(k, 0):8d}" for k in ["_total"] + KEYS) + "  " + name[:110])
print("  " + " ".join(f"{tot.get(k, 0):8d}" for k in ["_total"] + KEYS) + "  ALL KERNELS")
