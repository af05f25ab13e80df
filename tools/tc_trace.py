#!/usr/bin/env python
"""Dump the event trace of CTA 0 of one tcgen05 dilated-conv launch (tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import net as onet
from ubdvss_b200.engine import Engine
from ubdvss_b200 import _lib

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = Engine(precision="tf32")
eng.set_weights(onet.init_weights(0, seed=1234))
x = np.maximum(np.random.default_rng(0).normal(0, 1, size=(n, 256, 256, 24)), 0).astype(np.float32)
eng.debug_dilated_layer(x, layer, "tf32")          # warm-up (weights image, allocations)
eng.set_option("tc_trace", 1)
eng.debug_dilated_layer(x, layer, "tf32")
tr = np.zeros((3, 1024, 4), np.int64)
_lib.check(eng.handle, eng._lib.ubd_debug_read_trace(eng.handle, _lib.ptr(tr), tr.size))
t0 = min(tr[r, 0, 0] for r in range(3) if tr[r, 0, 0] > 0)
names = ["producer: wait_start wait_end issued", "mma: row_start seg_start tempty_ok issued", "epilogue(w2): wait_start tfull_ok stored"]
for r in range(3):
    ev = tr[r]
    k = int((ev[:, 0] > 0).sum())
    print(f"== role {r} ({names[r]}), {k} events")
    for i in range(min(k, 44)):
        print(i, [int(v - t0) if v else 0 for v in ev[i]])
    if k > 2:
        d = np.diff(ev[:k, 0])
        print("  inter-event cycles: median", int(np.median(d)), "mean", int(d.mean()), "total", int(ev[k - 1, 0] - ev[0, 0]))
