#!/usr/bin/env python
"""torchrun check of the data-parallel training exchange: the library's own NCCL all-reduce (ubd_comm_init /
ubd_allreduce_grads) against the torch.distributed all-reduce of the same gradient buffer; replicas must end up
with identical weights.  Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from ubdvss_b200 import losses, synth
from ubdvss_b200.net import Adam, B200Model, NetConfig

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
w0 = synth.synth_weights(0, seed=5)
x = synth.synth_images(4, 128, 192, seed=10 + rank)
y = synth.synth_targets(4, 32, 48, 0, seed=10 + rank)
res = {}
for native in (True, False):
    m = B200Model(NetConfig(), device=local, weights=w0)
    m.compile(Adam(1e-3), loss=losses.get_loss(False))
    m.set_distributed(True, native=native)
    for _ in range(3):
        out = m.train_on_batch(x, y, preprocessing="mobilenet_like")
    res[native] = np.concatenate([a.ravel() for a in m.get_weights()])
# the exchanged gradient of the CUDA path against the oracle: mean over the ranks of the per-replica gradients
m = B200Model(NetConfig(), device=local, weights=w0)
m.compile(Adam(1e-3), loss=losses.get_loss(False))
m.set_distributed(True, native=True)
m._engine.train_step(x, y, 1)
scale = m._allreduce_grads()
g_mine = np.concatenate([a.ravel() for a in m._engine.get_grads()]) * scale
if rank == 0:
    from oracle import loss as L, net as onet
    per = []
    for r in range(world):
        xr = onet.preprocess(synth.synth_images(4, 128, 192, seed=10 + r).astype(np.float64), "mobilenet_like").astype(np.float32)
        yr = synth.synth_targets(4, 32, 48, 0, seed=10 + r)
        per.append(np.concatenate([a.ravel() for a in L.train_step_torch(w0, xr, yr, False)[2]]))
    g_ref = np.mean(per, axis=0)
    gerr = float(np.abs(g_mine - g_ref).max() / np.abs(g_ref).max())
    print(f"all-reduced gradient vs oracle mean of per-replica gradients: rel err {gerr:.3e}")
    assert gerr <= 2e-3
diff = float(np.abs(res[True] - res[False]).max())
t = torch.from_numpy(res[True]).cuda()
lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
spread = float((hi - lo).abs().max())
moved = float(np.abs(res[True] - np.concatenate([a.ravel() for a in w0])).max())
if rank == 0:
    print(f"world {world}: |native - torch| = {diff:.3e}, replica spread = {spread:.3e}, weights moved by {moved:.3e}")
    assert diff <= 1e-6 and spread == 0.0 and moved > 0
    print("dist train check ok")
dist.barrier(); dist.destroy_process_group()
