"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
items = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":      # a CUDA line summary row
        ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        try:
            items.append((float(r[ie] or 0), float(r[ss] or 0), cur, r[0], r[1].strip()))
        except ValueError:
            pass
tot = sum(i[0] for i in items) or 1
tots = sum(i[1] for i in items) or 1
print("total warp-instructions %.0f, stall samples %.0f" % (tot, tots))
items.sort(key=lambda t: -t[0])
for it in items[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * it[0] / tot, 100 * it[1] / tots, it[2], it[3], it[4][:100]))
