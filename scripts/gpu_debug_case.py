import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import __graft_entry__ as ge
ge.load_package()
po = ge.load_oracle()
import helpers
from h264_scroll_encoder_b200 import abi, workloads, binding
gc = binding.GpuComposer(0); orc = po.Oracle()
seed=3; nr=5
b = workloads.make_random(5 + 9 * seed, 4 + 5 * seed, 3, seed=seed, num_refs=nr, flags=0)
hints = helpers.dense_hints(po, b)
frames, sizes = gc.compose_host(b.desc, hints, b.dyn, b.dyn_tab, b.payload)
for f in range(b.n_frames):
    want = helpers.checker_frame(orc, b, hints, f); got = frames[f]
    m = min(len(got), len(want)); d = np.nonzero(got[:m] != want[:m])[0]
    print("frame", f, len(got), len(want), "ndiff", len(d), "first", d[:10])
    if len(d):
        k = int(d[0]); print(" got ", got[k-4:k+12]); print(" want", want[k-4:k+12])
        # which MB: stage info
        h = hints[f*b.hint_stride:f*b.hint_stride+b.n_mbs]
        print(" dyn mbs:", int((h['kind']!=0).sum()), "W,H", b.mb_w, b.mb_h)
