"""Per-source-line attribution of an ncu capture (compiled with -lineinfo, captured with --import-source on):
    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > page.csv ; python tools/ncu_lines.py page.csv [top]
prints the lines with the most warp-stall samples (share of samples, share of executed warp instructions)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr, data, cur = None, [], None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        data.append((cur, r))
ci = hdr.index("Warp Stall Sampling (All Samples)")
ii = hdr.index("Instructions Executed")
tot = sum(int(r[ci]) for _, r in data)
toti = sum(int(r[ii]) for _, r in data)
print("total samples", tot, "warp instructions", toti)
for f, r in sorted(data, key=lambda fr: -int(fr[1][ci]))[:top]:
    print("%5.2f%% smp %5.2f%% ins  %s:%s  %s" % (100 * int(r[ci]) / tot, 100 * int(r[ii]) / toti, f.split("/")[-1], r[0], r[1].strip()[:100]))
