"""CPU (gloo, world_size 2) coverage of the multi-process host logic: batch sharding for inference
and the gradient all-reduce + 1/world scaling of data-parallel training."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ubdvss_b200.parallel import allreduce_mean_, shard_bounds


def test_shard_bounds_cover_batch():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import loss as L, net as onet
        from ubdvss_b200 import synth
        torch.set_num_threads(1)
        w = onet.init_weights(0, seed=7)
        x = synth.synth_images(4, 64, 64, seed=5).astype(np.float32) / 127.5 - 1
        y = synth.synth_targets(4, 16, 16, 0, seed=5)
        lo, hi = shard_bounds(4, world, rank)
        _, _, grads, _ = L.train_step_torch(w, x[lo:hi], y[lo:hi], False)       # per-replica loss
        flat = torch.from_numpy(np.concatenate([g.ravel() for g in grads]).astype(np.float32))
        scale = allreduce_mean_(flat, dist)
        if rank == 0:
            out.put((flat.numpy() * scale, scale))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_is_mean_of_replica_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, scale = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from oracle import loss as L, net as onet
    from ubdvss_b200 import synth
    w = onet.init_weights(0, seed=7)
    x = synth.synth_images(4, 64, 64, seed=5).astype(np.float32) / 127.5 - 1
    y = synth.synth_targets(4, 16, 16, 0, seed=5)
    per = []
    for r in range(2):
        lo, hi = shard_bounds(4, 2, r)
        g = L.train_step_torch(w, x[lo:hi], y[lo:hi], False)[2]
        per.append(np.concatenate([a.ravel() for a in g]))
    assert scale == 0.5
    assert np.allclose(got, (per[0] + per[1]) / 2, rtol=1e-5, atol=1e-7)
