"""GPU box: raw pinned-memory PCIe bandwidth, one direction and both at once, for several copy granularities
(the floor of the end-to-end path)."""
import torch, time
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, piece_in, piece_out, reps=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                for o in range(0, n, piece_in): d_in[o:o + piece_in].copy_(h_in[o:o + piece_in], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                for o in range(0, n, piece_out): h_out[o:o + piece_out].copy_(d_out[o:o + piece_out], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps
for pin, pout in ((n, n), (32 << 20, 32 << 20), (4 << 20, 32 << 20), (1 << 20, 32 << 20), (256 << 10, 32 << 20), (32 << 20, 4 << 20)):
    for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
        run(a, b, pin, pout, 1); dt = run(a, b, pin, pout)
        print("pieces in %6d KB out %6d KB  %4s: %.2f ms per 256 MiB -> %.1f GB/s per direction" % (pin >> 10, pout >> 10, name, dt * 1e3, n / dt / 1e9))
