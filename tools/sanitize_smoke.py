#!/usr/bin/env python
"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import net as onet
from ubdvss_b200 import _lib, synth
from ubdvss_b200.engine import Engine

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
eng = Engine(precision=prec, n_classes=2)
eng.set_weights(onet.init_weights(2, seed=1))
x = synth.synth_images(2, 64, 320, seed=1)
lg = eng.forward(x[:1], _lib.PREPROC_MOBILENET)
thr = float(np.quantile(lg[..., 0], 0.8))
mask, logits, labels, comps, counts = eng.segment(x, thr, 10, _lib.PREPROC_MOBILENET, want_labels=True)
print(prec, "segment ok", counts.tolist(), float(mask.mean()))
# multi-row pieces per CTA, two strips per row in the stem (W/2 = 512 > 256)
x2 = synth.synth_images(3, 320, 1024, seed=2)
mask2, _, _, _, counts2 = eng.segment(x2, thr, 10, _lib.PREPROC_MOBILENET)
print(prec, "segment (large) ok", counts2.tolist())
xf = (x.astype(np.float32) - 127.5) / 127.5
eng.forward(xf, _lib.PREPROC_NONE)
m = synth.stress_masks(3, 40, 72, seed=2)
eng.postprocess(m, None, 10, want_labels=True)
if prec in ("fp32", "tf32"):     # tf32: tcgen05 forward / backward-data + the mma.sync weight-gradient kernels
    y = synth.synth_targets(2, 16, 80, 2, seed=1)
    p = eng.train_step(x, y, _lib.PREPROC_MOBILENET); eng.adam_step()
    print("train ok", p.tolist())
    y3 = synth.synth_targets(3, 80, 256, 2, seed=2)       # several rows per CTA, two strips per half-resolution row
    p = eng.train_step(x2, y3, _lib.PREPROC_MOBILENET)
    print("train (large) ok", p.tolist())
# pipelined calls: three batches in flight, CC on its own stream
ts = [eng.segment_submit(x, thr, 10, _lib.PREPROC_MOBILENET) for _ in range(3)]
for t in ts:
    eng.segment_wait(t)
print(prec, "pipelined ok")
