"""Bring-up helper for the tcgen05 dilated-layer kernels: one layer, one shape, against the oracle."""
import sys, time
import numpy as np
from oracle import net as onet
from ubdvss_b200.engine import Engine

shape = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (12, 64, 64)
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
e = Engine()
w = onet.init_weights(0, seed=1234)
e.set_weights(w)
rng = np.random.default_rng(0)
x = onet.round_tf32(np.maximum(rng.normal(0, 1, size=shape + (24,)), 0).astype(np.float32))
if prec == "bf16":
    u = x.view(np.uint32); x = ((u + np.uint32(0x7FFF) + ((u >> 16) & 1)) & np.uint32(0xFFFF0000)).view(np.float32)
t0 = time.time()
try:
    got = e.debug_dilated_layer(x, layer, prec)
except Exception as ex:
    print("ERR", ex, "after", time.time() - t0); sys.exit(0)
print("ran in", time.time() - t0)
k, b = w[9 + 2 * layer], w[10 + 2 * layer]
if prec == "tf32":
    k = onet.round_tf32(k)
else:
    u = np.ascontiguousarray(k).view(np.uint32); k = ((u + np.uint32(0x7FFF) + ((u >> 16) & 1)) & np.uint32(0xFFFF0000)).view(np.float32)
ref = np.maximum(onet._conv3x3_np(x.astype(np.float64), k.astype(np.float64), onet.DILATIONS[layer]) + b, 0)
err = np.abs(got - ref)
print("max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
bad = np.argwhere(err.max(axis=-1) > 1e-3)
print("bad px", len(bad), "rows:", sorted(set((int(a), int(b_)) for a, b_, _ in bad))[:40])
