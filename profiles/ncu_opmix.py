"""Opcode mix of the executed warp-instructions, overall and per pipeline phase (source-line ranges as in ncu_phases.py).
Usage: ncu_opmix.py src.csv [n_phases]   (src.csv = ncu -i rep --page source --print-source cuda,sass --csv)"""
import csv, sys, re, collections, os
rows = list(csv.reader(open(sys.argv[1])))
src = {}
def lines_of(fn):
    for base in ("h264-scroll-encoder_b200/csrc/", ""):
        if os.path.exists(base + fn): return open(base + fn).read().split("\n")
    return []
def phase(fn, ln):
    if fn not in src: src[fn] = lines_of(fn)
    L = src[fn]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.search(r"/\* -{2,} (.+?)( -{2,}|$)", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(1).strip()[:40]
        m = re.match(r"^(HSD|__device__|__global__|static|template).*?(\w+)\s*\(", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(2)
    return fn
cur = None; hdr = None; seen = set(); ln = 0
tot = collections.Counter(); per = collections.defaultdict(collections.Counter)
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        try: ln = int(r[0])
        except ValueError: pass
        continue
    if hdr and len(r) == len(hdr) and r[2] != "":
        if (cur, r[2]) in seen: continue
        seen.add((cur, r[2]))
        try: n = float(r[7] or 0)
        except ValueError: continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
        if not m: continue
        op = m.group(2)
        tot[op] += n; per[phase(cur, ln)][op] += n
T = sum(tot.values()) or 1
CTRL = ("BRA", "BSSY", "BSYNC", "ISETP", "PLOP3", "UISETP", "WARPSYNC", "BREAK", "EXIT", "CALL", "RET")
print("total %.0f warp-instructions; control (branches, reconvergence, predicates) %.1f%%" % (T, 100 * sum(tot[o] for o in CTRL) / T))
print("  " + "  ".join("%s %.1f" % (o, 100 * n / T) for o, n in tot.most_common(24)))
for ph, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[: int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    n = sum(c.values())
    print("%5.1f%%  %-52s ctrl %4.1f%%  %s" % (100 * n / T, ph, 100 * sum(c[o] for o in CTRL) / n, " ".join("%s %.0f" % (o, 100 * v / n) for o, v in c.most_common(7))))
