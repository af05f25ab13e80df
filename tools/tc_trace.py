#!/usr/bin/env python
"""Dump the event trace of CTA 0 of one tcgen05 dilated-conv launch (tuning aid).
Needs a trace-enabled build: UBD_TC_TRACE=1 python -m ubdvss_b200.build --force."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import net as onet
from ubdvss_b200.engine import Engine
from ubdvss_b200 import _lib

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
eng = Engine(precision=prec)
eng.set_weights(onet.init_weights(0, seed=1234))
x = np.maximum(np.random.default_rng(0).normal(0, 1, size=(n, 256, 256, 24)), 0).astype(np.float32)
eng.debug_dilated_layer(x, layer, prec)          # warm-up (weights image, allocations)
if len(sys.argv) > 4: eng.set_option('tc_variant', int(sys.argv[4]))
eng.set_option("tc_trace", 1)
eng.debug_dilated_layer(x, layer, prec)
tr = np.zeros((4, 1024, 4), np.int64)
_lib.check(eng.handle, eng._lib.ubd_debug_read_trace(eng.handle, _lib.ptr(tr), tr.size))
t0 = min(tr[r, 0, 0] for r in range(4) if tr[r, 0, 0] > 0)
names = ["producer: wait_start wait_end issued", "mma: row_start gempty_ok full_ok issued", "epilogue seg0 q0: wait_start gfull_ok rearmed stored", "-"]
for r in range(4):
    ev = tr[r]
    k = int((ev[:, 0] > 0).sum())
    print(f"== role {r} ({names[r]}), {k} events")
    for i in range(min(k, 44)):
        print(i, [int(v - t0) if v else 0 for v in ev[i]])
    if k > 2:
        d = np.diff(ev[:k, 0])
        print("  inter-event cycles: median", int(np.median(d)), "mean", int(d.mean()), "total", int(ev[k - 1, 0] - ev[0, 0]))
