"""Bring-up probe of the tensor-core weight-gradient kernel (ubd_wgrad.cuh): delta inputs show which (ky, kx, ic, oc) an
(x pixel, g pixel) pair lands on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import net as onet
from ubdvss_b200.engine import Engine

eng = Engine(precision="tf32")
eng.set_weights(onet.init_weights(0, seed=1))
rng = np.random.default_rng(0)
shape = (1, 16, 32, 24)
x = onet.round_tf32(np.maximum(rng.normal(0, 1, size=shape), 0).astype(np.float32))
g = onet.round_tf32(rng.normal(0, 1, size=shape).astype(np.float32))
dk, db = eng.debug_wgrad(x, g, 1)
print("random: |dk| max", np.abs(dk).max(), "nonzero", np.count_nonzero(dk), "db", db[:6], "ref db", g.sum((0, 1, 2))[:6])
for (yx, xx, ic), (yg, xg, oc), d in [((5, 9, 3), (5, 9, 7), 1), ((5, 9, 3), (5, 10, 7), 1), ((5, 9, 3), (6, 9, 7), 1), ((5, 9, 13), (5, 9, 21), 1),
                                       ((5, 9, 3), (3, 11, 7), 2), ((0, 0, 0), (0, 0, 0), 1)]:
    x = np.zeros(shape, np.float32); g = np.zeros(shape, np.float32)
    x[0, yx, xx, ic] = 1.0; g[0, yg, xg, oc] = 2.0
    dk, db = eng.debug_wgrad(x, g, d)
    nz = np.argwhere(dk != 0)
    print(f"x@{(yx, xx, ic)} g@{(yg, xg, oc)} d={d}: nonzero dk {[(tuple(i), float(dk[tuple(i)])) for i in nz[:8]]} db nz {[(int(i), float(db[i])) for i in np.nonzero(db)[0]]}")
