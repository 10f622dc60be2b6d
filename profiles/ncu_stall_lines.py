"""Per source line: which stall reasons its samples carry.  Usage: ncu_stall_lines.py src.csv [reason ...]
(src.csv = ncu -i rep --page source --print-source cuda,sass --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
reasons = sys.argv[2:] or ["stall_long_sb", "stall_no_inst", "stall_barrier"]
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        key = (cur, int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, {})
        for k in reasons + ["Warp Stall Sampling (All Samples)", "Instructions Executed"]:
            try: a[k] = a.get(k, 0) + float(r[hdr.index(k)] or 0)
            except ValueError: pass
tot = {k: sum(a.get(k, 0) for a in agg.values()) or 1 for k in reasons + ["Warp Stall Sampling (All Samples)"]}
print("totals:", {k: int(v) for k, v in tot.items()})
for k in reasons:
    print("---", k)
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1].get(k, 0))[:14]:
        print("%5.1f%%  %s:%d  %s" % (100 * a.get(k, 0) / tot[k], key[0], key[1], key[2]))
