"""SASS instruction count per pipeline phase of one kernel (code size, not execution counts).
Usage: sass_phases.py <nvdisasm -g -c output> <section substring>"""
import re, sys, collections, os
dis, want = sys.argv[1], sys.argv[2]
src = {}
def lines_of(fn):
    for base in ("h264-scroll-encoder_b200/csrc/", ""):
        p = base + fn
        if os.path.exists(p): return open(p).read().split("\n")
    return []
def phase(fn, ln):
    if fn not in src: src[fn] = lines_of(fn)
    L = src[fn]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.search(r"/\* -{2,} (.+?)( -{2,}|$)", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(1).strip()[:48]
        m = re.match(r"^(HSD|__device__|__global__|static|template).*?(\w+)\s*\(", L[i])
        if m: return fn.split(".")[0][:8] + ":" + m.group(2)
    return fn
fn = None; cur = None; counts = collections.Counter(); per_fn = collections.Counter()
for line in open(dis):
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
    if m: fn = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and fn and want in fn:
        per_fn[fn[:60]] += 1
        counts[phase(*cur) if cur else "?"] += 1
print("sections:", dict(per_fn))
tot = sum(counts.values())
print("total SASS instructions", tot, "= %.1f KB" % (tot * 16 / 1024))
for k, v in counts.most_common(60): print("%5d %5.1f%%  %s" % (v, 100 * v / tot, k))
