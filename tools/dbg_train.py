"""Per-array gradient error of the tf32 (tensor-core) training step against the autograd oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import loss as L, net as onet
from ubdvss_b200 import _lib, synth
from ubdvss_b200.engine import Engine

shape = (2, 64, 96)
w = onet.init_weights(0, seed=3)
n, H, W = shape
x = synth.synth_images(n, H, W, seed=3)
y = synth.synth_targets(n, H // 4, W // 4, 0, seed=3)
xf = onet.preprocess(x.astype(np.float64), "mobilenet_like").astype(np.float32)
loss, _, ref, _ = L.train_step_torch(w, xf, y, False)
for tc in (1, 2, 4, 7):
    eng = Engine(precision="tf32")
    eng.set_weights(w)
    eng.set_option("train_tc_bits", tc)
    parts = eng.train_step(x, y, _lib.PREPROC_MOBILENET)
    print("train_tc", tc, "loss", parts[0], "ref", loss)
    for i, (g, r) in enumerate(zip(eng.get_grads(), ref)):
        print(f"  {i:2d} {str(g.shape):18s} max|r| {np.abs(r).max():.3e} max err {np.abs(g - r).max():.3e} rel-l2 {np.linalg.norm((g - r).ravel()) / (np.linalg.norm(r.ravel()) + 1e-30):.3e}")
