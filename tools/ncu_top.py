"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions with their stall reasons."""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
h = rows[1]
isrc, iss, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = []
for idx, r in enumerate(rows[2:]):
    if len(r) <= iss:
        continue
    try:
        s = int(r[iss])
    except ValueError:
        continue
    data.append((s, idx, r))
tot = sum(s for s, _, _ in data)
print("total samples", tot, "instruction rows", len(data))
for s, idx, r in sorted(data, key=lambda x: -x[0])[:topn]:
    st = {h[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{s:7d} {100 * s / tot:5.1f}% #{idx:5d} ex={r[iex]:>9s} {r[isrc][:64]:64s} {top}")
