"""-m gpu: data-parallel training exchange of the CUDA path.  With >= 2 GPUs the NCCL path is launched
under torchrun (tools/dist_train_check.py: the library's own communicator, ubd_comm_init / ubd_allreduce_grads, against
torch.distributed and against the oracle's mean of per-replica gradients); on a single GPU two handles play the two
ranks and the sum of their flat gradient buffers stands in for the all-reduce."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    from ubdvss_b200 import _lib
    return _lib.load().ubd_device_count()


def test_two_replicas_on_one_gpu_match_oracle_mean_gradient():
    import torch
    from oracle import loss as L, net as onet
    from ubdvss_b200 import losses, synth
    from ubdvss_b200.net import Adam, B200Model, NetConfig
    w0 = synth.synth_weights(0, seed=5, calibrated=True)
    xs = [synth.synth_images(4, 128, 192, seed=10 + r) for r in range(2)]
    ys = [synth.synth_targets(4, 32, 48, 0, seed=10 + r) for r in range(2)]
    reps = []
    for r in range(2):
        m = B200Model(NetConfig(), weights=w0)
        m.compile(Adam(1e-3), loss=losses.get_loss(False))
        m._engine.train_step(xs[r], ys[r], 1)
        reps.append(m)

    def view(m):
        ptr, n = m._engine.grad_buffer()

        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(d, device="cuda:0")
    for m in reps:
        m._engine.synchronize()
    g = [view(m) for m in reps]
    total = g[0] + g[1]                       # what ncclAllReduce(sum) leaves on every rank
    for t in g:
        t.copy_(total)
    torch.cuda.synchronize()
    per = []
    for r in range(2):
        xr = onet.preprocess(xs[r].astype(np.float64), "mobilenet_like").astype(np.float32)
        per.append(np.concatenate([a.ravel() for a in L.train_step_torch(w0, xr, ys[r], False)[2]]))
    ref = np.mean(per, axis=0)
    mine = np.concatenate([a.ravel() for a in reps[0]._engine.get_grads()]) * 0.5
    assert np.abs(mine - ref).max() <= 2e-3 * np.abs(ref).max()
    for m in reps:
        o = m._optimizer
        m._engine.adam_step(o.lr, o.beta_1, o.beta_2, o.epsilon, 0.5)
    wa, wb = (np.concatenate([a.ravel() for a in m.get_weights()]) for m in reps)
    assert np.array_equal(wa, wb) and np.abs(wa - np.concatenate([a.ravel() for a in w0])).max() > 0


@pytest.mark.skipif("_n_gpus() < 2")
def test_nccl_exchange_under_torchrun():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tools", "dist_train_check.py")],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "dist train check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
