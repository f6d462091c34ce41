"""Top stall lines of an `ncu --page source --csv` dump: python tools/ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[idx["# Samples"]]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot)
for i, (k, r) in enumerate(data):
    r.append(i)
for k, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((int(r[idx[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{k:6d} {100*k/tot:5.1f}% #{r[-1]:5d} {r[idx['Source']][:90]:90s} {top}")
