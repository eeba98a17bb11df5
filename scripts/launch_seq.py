"""Development tool: print the kernel sequence of one scan from an ncu launch list (gpu__time_duration.sum, csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else -3
hdr, seq = None, []
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            k = d["Kernel Name"].split("(")[0]
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
            seq.append((k, v, d.get("Grid Size", ""), d.get("Block Size", "")))
idx = [i for i, s in enumerate(seq) if s[0].startswith("k_deskew_var_init")]
i0, i1 = idx[which], idx[which + 1]
tot = 0.0
for k, v, g, b in seq[i0:i1]:
    print(f"{k[:56]:56s} {v:8.2f}  {g} {b}")
    if "elementwise" not in k:
        tot += v
print("sum (without the L2 flush)", round(tot, 1))
