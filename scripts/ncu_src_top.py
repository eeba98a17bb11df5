"""Development tool: per-source-line stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_src_top.py report.ncu-rep kernel_regex [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
# one block per (launch, source file): "File Path" line, "Function Name" line, header, rows
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"File Path"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur:
    blocks.append(cur)
files = sorted(set(b[0] for b in blocks))
per_launch = len(files)
print(len(blocks), "blocks,", per_launch, "source files per launch")
rows = []
hdr = None
for b in blocks[which * per_launch:(which + 1) * per_launch]:
    rr = list(csv.reader(io.StringIO("\n".join(b[2:]))))
    hdr = rr[0]
    fn = b[0].split(",")[1].strip('"').split("/")[-1]
    for r in rr[1:]:
        if len(r) == len(hdr) and r[0] != "":  # source-line rows only (SASS rows have an empty line number)
            r[0] = fn[:12] + ":" + r[0]
            rows.append(r)
rows = [hdr] + rows
iS = hdr.index("# Samples")
iSrc = 1
iI = hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in rows[1:] if len(r) == len(hdr))
print("total samples", tot, "columns:", hdr[:4])
rs = [r for r in rows[1:] if len(r) == len(hdr) and int(r[iS] or 0) > 0]
rs.sort(key=lambda r: -int(r[iS]))
for r in rs[:top]:
    stalls = sorted(((int(r[i] or 0), hdr[i][6:]) for i in st), reverse=True)[:3]
    print(f"{int(r[iS]):6d} {100*int(r[iS])/tot:5.1f}%  inst {r[iI]:>8s}  {r[0][:18]:>18s} {r[iSrc].strip()[:100]:100s} {stalls}")
