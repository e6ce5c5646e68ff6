"""Hottest SASS lines of a kernel in an .ncu-rep by warp-stall samples: python tools/ncu_hot.py rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iall, ins = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= iall:
        continue
    try:
        n = int(r[ins] or 0)
    except ValueError:
        continue
    data.append((n, r))
total = sum(n for n, _ in data)
print(f"total samples {total}")
for n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{100 * n / max(total, 1):5.1f}%  {r[isrc][:90]:90s}  {[(b, a) for a, b in st if a]}")
