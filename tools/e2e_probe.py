#!/usr/bin/env python
"""Where the end-to-end time of ubd_segment goes: device-resident call vs host-buffer call variants."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import net as onet
from ubdvss_b200 import _lib, synth
from ubdvss_b200.engine import Engine

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, S = 64, 1024
eng = Engine(precision=prec)
eng.set_weights(onet.init_weights(0, seed=1234))
imgs = np.ascontiguousarray(np.concatenate([synth.synth_images(8, S, S, seed=1)] * 8, 0))
pinned = torch.from_numpy(imgs).pin_memory(); h_imgs = pinned.numpy()
pageable = imgs.copy()
d_imgs = pinned.cuda()
thr = float(np.quantile(eng.forward(imgs[:2], _lib.PREPROC_MOBILENET)[..., 0], 0.9))
mask_h = torch.empty((B, S // 4, S // 4), dtype=torch.uint8).pin_memory().numpy()
d_mask = torch.empty((B, S // 4, S // 4), dtype=torch.uint8, device="cuda")
cap = 64 * B
comps_h = np.zeros(cap, _lib.COMPONENT_DTYPE); counts_h = np.zeros(B, np.int32)

def run(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{name:50s} {dt:7.3f} ms  {B / dt * 1e3:9.0f} img/s")

def seg(x, mask):
    _lib.check(eng.handle, eng._lib.ubd_segment(eng.handle, _lib.ptr(x), _lib.UBD_U8, B, S, S, _lib.PREPROC_MOBILENET,
                                                 np.float32(thr), 10, _lib.ptr(mask) if mask is not None else None, None, None,
                                                 _lib.ptr(comps_h), cap, _lib.ptr(counts_h)))
run("segment_dev (device in, device mask)", lambda: eng.segment_dev(d_imgs.data_ptr(), _lib.UBD_U8, B, S, S, thr, 10, _lib.PREPROC_MOBILENET, d_mask.data_ptr(), 0))
run("ubd_segment pinned in, no mask out", lambda: seg(h_imgs, None))
run("ubd_segment pinned in, pinned mask out", lambda: seg(h_imgs, mask_h))
run("ubd_segment pageable in, pinned mask out", lambda: seg(pageable, mask_h))
t0 = time.perf_counter()
for _ in range(20): pinned.cuda(non_blocking=True)
torch.cuda.synchronize(); print("H2D 64 MiB pinned:", (time.perf_counter() - t0) / 20 * 1e3, "ms")
