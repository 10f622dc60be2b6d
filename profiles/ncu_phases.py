"""Group an `ncu --page source --print-source cuda,sass --csv` dump into pipeline phases (by source line ranges)."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None; items = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        ie, ss = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        try: items.append((float(r[ie] or 0), float(r[ss] or 0), cur, int(r[0]), r[1].strip()))
        except ValueError: pass
src = {}
def lines_of(fn):
    import os
    for base in ("h264-scroll-encoder_b200/csrc/", ""):
        p = base + fn
        if os.path.exists(p): return open(p).read().split("\n")
    return []
def phase(fn, ln, text):
    if fn not in src: src[fn] = lines_of(fn)
    L = src[fn]
    # walk back to the nearest marker comment "/* ---- name" or function header
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.search(r"/\* -{2,} (.+?)( -{2,}|$)", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(1).strip()[:40]
        m = re.match(r"^(HSD|__device__|__global__|static|template).*?(\w+)\s*\(", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(2)
    return fn
agg = {}
for ie, ss, fn, ln, text in items:
    k = phase(fn, ln, text)
    a = agg.setdefault(k, [0, 0]); a[0] += ie; a[1] += ss
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-instructions %.0f, stall samples %.0f" % (ti, ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5.1f%% inst %5.1f%% stall  %s" % (100 * a[0] / ti, 100 * a[1] / ts, k))
