"""Per-warp event trace of one CTA (COMPOSE_TRACE=1 build, bench.py writes gpurun_out/trace.bin): for every event, when the
first and the last warp got there.  Events: 1 frame start; 10/11 arrive at / leave the tile's first barrier; 12 decide done
(arrive at the offsets scan); 13 scan done; 14 concatenation + merge done; 15/16 push done, barrier; 17 splice done;
18/19 barrier in front of emit; 30/31 classify; 32/33 masks; 34 counts scanned; 35/36 clean stores; 37 listed stores;
20/21 barrier behind emit.   Usage: trace_view.py trace.bin [n_warps]"""
import sys, numpy as np
t = np.fromfile(sys.argv[1], np.uint64).reshape(32, 1024)
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ids = (t >> np.uint64(48)).astype(np.int64); clk = (t & np.uint64((1 << 48) - 1)).astype(np.int64)
for w in range(32):                                      # (events 40+ are per-warp details of the splice loop: not part of the common sequence)
    keep = np.nonzero((ids[w] != 0) & (ids[w] < 40))[0]
    i2, c2 = np.zeros(1024, np.int64), np.zeros(1024, np.int64)
    i2[:len(keep)] = ids[w, keep]; c2[:len(keep)] = clk[w, keep]
    ids[w], clk[w] = i2, c2
n = min(int((ids[w] != 0).sum()) for w in range(nw))
print("events per warp:", [int((ids[w] != 0).sum()) for w in range(nw)])
# align on the first frame-start event every warp has
starts = [list(np.nonzero(ids[w, :n] == 1)[0]) for w in range(nw)]
k = min(len(s) for s in starts)
if k < 2: sys.exit("no complete frame in the trace")
names = {1: "frame start", 10: "tile: arrive", 11: "tile: go", 12: "decide done", 13: "scan done", 14: "write done", 15: "push done", 16: "push barrier", 17: "splice done",
         18: "emit: arrive", 19: "emit: go", 30: "classify done", 31: "classify barrier", 32: "masks done", 33: "masks barrier", 34: "counts scanned",
         35: "clean stores done", 36: "stores barrier", 37: "listed stores done", 20: "emit returned", 21: "frame barrier"}
for fr in range(1, k - 1):
    ref = clk[0, starts[0][fr]]                      # the same frame in every warp: the frame start closest in time to warp 0's
    sel = [min(range(len(starts[w]) - 1), key=lambda j: abs(clk[w, starts[w][j]] - ref)) for w in range(nw)]
    a = [starts[w][sel[w]] for w in range(nw)]; b = [starts[w][sel[w] + 1] for w in range(nw)]
    m = min(b[w] - a[w] for w in range(nw))
    t0 = min(clk[w, a[w]] for w in range(nw))
    print("frame %d: %d events, %.1f us at 1.9 GHz" % (fr, m, (min(clk[w, b[w]] for w in range(nw)) - t0) / 1900.0))
    print("   id  event                first    last   spread  last warp   work(min/avg/max since the previous event)")
    for i in range(m):
        e = [ids[w, a[w] + i] for w in range(nw)]
        c = np.array([clk[w, a[w] + i] for w in range(nw)]) - t0
        p = np.array([clk[w, a[w] + i - 1] for w in range(nw)]) - t0 if i else c
        d = c - p
        tag = names.get(e[0], "?") if len(set(e)) == 1 else "MIXED %s" % sorted(set(e))
        print("  %3d  %-18s %7d %7d  %6d   w%-2d        %6d %6d %6d" % (e[0], tag, c.min(), c.max(), c.max() - c.min(), int(c.argmax()), d.min(), int(d.mean()), d.max()))
    if len(sys.argv) > 3:                                # per-warp work for the listed event ids
        want = [int(x) for x in sys.argv[3].split(",")]
        for i in range(m):
            e = ids[0, a[0] + i]
            if e in want:
                d = [int(clk[w, a[w] + i] - clk[w, a[w] + i - 1]) for w in range(nw)]
                print("   event %d (#%d) work per warp: %s" % (e, i, " ".join("%d" % x for x in d)))
