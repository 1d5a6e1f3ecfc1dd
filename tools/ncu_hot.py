"""Hottest SASS instructions (warp-stall samples) of an .ncu-rep: python tools/ncu_hot.py rep [top]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        print("==", rows[i][1][:100])
        hdr = rows[i + 1]
        si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        body = []
        j = i + 2
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            body.append(rows[j]); j += 1
        tot = sum(float(r[si] or 0) for r in body) or 1
        order = sorted(range(len(body)), key=lambda k: -float(body[k][si] or 0))[:top]
        for k in sorted(order):
            r = body[k]
            print(f"{float(r[si])/tot*100:5.1f}%  #{k:4d} x{r[ii]:>9s}  {r[1].strip()[:100]}")
        i = j
    else:
        i += 1
