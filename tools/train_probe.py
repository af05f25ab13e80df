"""Config E probe: a few training steps (32 x 512x512, forward + loss + backward + Adam) through B200Model.train_on_batch.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list, or plainly for the step time."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ubdvss_b200 import losses as ulosses, synth                      # noqa: E402
from ubdvss_b200.net import Adam, B200Model, NetConfig                # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
tb = 32
tx = np.concatenate([synth.synth_images(8, 512, 512, seed=40)] * (tb // 8))
ty = np.concatenate([synth.synth_targets(8, 128, 128, 0, seed=40)] * (tb // 8))
model = B200Model(NetConfig(), device=0, precision=precision, weights=synth.synth_weights(0, seed=1234, calibrated=True))
model.compile(Adam(1e-3), loss=ulosses.get_loss(False))
for _ in range(2):
    model.train_on_batch(tx, ty, preprocessing="mobilenet_like")
t0 = time.perf_counter()
for _ in range(steps):
    out = model.train_on_batch(tx, ty, preprocessing="mobilenet_like")
dt = (time.perf_counter() - t0) / steps
print(f"train step {dt * 1e3:.3f} ms  loss {out[0] if isinstance(out, (list, tuple)) else out}")
