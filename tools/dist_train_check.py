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
