#!/usr/bin/env python
"""Event trace of CTA 0 of the stem's L2 launch (L1-producer variant of the column-rotating kernel).
Needs a trace-enabled build: UBD_TC_TRACE=1 python -m ubdvss_b200.build --force."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import net as onet
from ubdvss_b200.engine import Engine
from ubdvss_b200 import _lib, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
eng = Engine(precision=prec)
eng.set_weights(onet.init_weights(0, seed=1234))
x = synth.synth_images(16, 1024, 1024, seed=3)
eng.forward(x, _lib.PREPROC_MOBILENET)
eng.set_option("tc_trace", 1)
eng.forward(x, _lib.PREPROC_MOBILENET)
tr = np.zeros((8, 1024, 4), np.int64)
_lib.check(eng.handle, eng._lib.ubd_debug_read_trace(eng.handle, _lib.ptr(tr), tr.size))
names = {4: "producer", 5: "mma seg0: row_start gempty_ok full_ok issued", 6: "epilogue seg0 q0: wait_start gfull_ok released stored",
         7: "L1 warp: row_start slot_free stored published"}
t0 = min(tr[r, 0, 0] for r in (5, 6, 7) if tr[r, 0, 0] > 0)
for r in (5, 6, 7):
    ev = tr[r]
    k = int((ev[:, 0] > 0).sum())
    print(f"== role {r} ({names[r]}), {k} events")
    for i in range(min(k, 40)):
        print(i, [int(v - t0) if v else 0 for v in ev[i]])
    if k > 2:
        dd = np.diff(ev[:k, 0])
        print("  inter-event cycles: median", int(np.median(dd)), "mean", int(dd.mean()), "total", int(ev[k - 1, 0] - ev[0, 0]))
