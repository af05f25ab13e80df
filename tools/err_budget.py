"""Error budget of the three precisions against a float64 run of the oracle (SURVEY 7 "TF32 tolerance"):
max |sigmoid(logit) - sigmoid(logit_fp64)| and logit errors on calibrated weights (logit std ~2, content-dependent)
and on plain Glorot weights, 2 x 1024^2 (+ class head C = 26).  gpurun: python tools/err_budget.py"""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import net as onet
from ubdvss_b200 import _lib, synth
from ubdvss_b200.engine import Engine

sig = lambda z: 1 / (1 + np.exp(-z.astype(np.float64)))
out = {}
for tag, n_classes, calibrated in (("calibrated_c0", 0, True), ("calibrated_c26", 26, True), ("glorot_c0", 0, False)):
    w = synth.synth_weights(n_classes, seed=1234, calibrated=calibrated)
    x = synth.synth_images(2, 1024, 1024, seed=5)
    xf = onet.preprocess(x.astype(np.float64), "mobilenet_like")
    ref64 = onet.forward_torch(w, xf, dtype=torch.float64)
    ref32 = onet.forward_torch(w, xf.astype(np.float32))
    row = {"logit_std": float(ref64[..., 0].std()), "oracle_fp32_vs_fp64_prob": float(np.abs(sig(ref32[..., 0]) - sig(ref64[..., 0])).max())}
    for prec in ("fp32", "tf32", "bf16"):
        e = Engine(precision=prec, n_classes=n_classes)
        e.set_weights(w)
        for variant in ((0, 2) if prec != "fp32" else (0,)):
            e.set_option("stem_variant", variant)
            got = e.forward(x, _lib.PREPROC_MOBILENET)
            key = prec + ("_fusedstem" if variant == 2 else "")
            row[key + "_prob"] = float(np.abs(sig(got[..., 0]) - sig(ref64[..., 0])).max())
            row[key + "_logit_max"] = float(np.abs(got - ref64).max())
            row[key + "_logit_rms"] = float(np.sqrt(np.mean((got - ref64) ** 2)))
            row[key + "_mask_flip_frac"] = float(np.mean((got[..., 0] > 0) != (ref64[..., 0] > 0)))
    out[tag] = row
print(json.dumps(out, indent=1))
