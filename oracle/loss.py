"""CPU restatement of the reference's pixelwise loss and optimizer step.  TEST INFRASTRUCTURE ONLY.

Follows ``semantic_segmentation/losses.py:13-17`` (weights), ``27-30`` (sigmoid / y_true>0),
``33-62`` (detection / combined loss), ``65-83`` (masked sparse softmax CE), ``86-126``
(positive / negative / hard-negative BCE) and ``train.py:110`` (``Adam(lr)``).
PARITY UNPINNED (TensorFlow-1.x / Keras-2.x absent, no reference tests); external semantics
encoded here are the documented ones:

* ``K.binary_crossentropy(target, output)`` on probabilities (TF backend): clip output to
  [eps, 1-eps] with eps = 1e-7 in float32, ``x = log(p/(1-p))``, then
  ``sigmoid_cross_entropy_with_logits = max(x,0) - x*t + log1p(exp(-|x|))`` (built with
  ``tf.where(x >= 0, ...)``, hence differentiable as ``sigmoid(x) - t`` also at x = 0).
* ``tf.clip_by_value`` passes the gradient where lo <= p <= hi (so saturated logits get 0).
* ``tf.nn.top_k`` keeps the lower index among equal values.
* Keras-2 ``Adam``: ``lr_t = lr*sqrt(1-b2^t)/(1-b1^t)``, ``p -= lr_t*m/(sqrt(v)+eps)``, eps=1e-7.
"""
from __future__ import annotations

import numpy as np

L_POSITIVE_WEIGHT = 15.0            # losses.py:13
L_NEGATIVE_WEIGHT = 1.0             # losses.py:14
L_HARD_NEGATIVE_WEIGHT = 5.0        # losses.py:15
L_DETECTION_WEIGHT = 1.0            # losses.py:16
L_CLASSIFICATION_WEIGHT = 1.0       # losses.py:17
KERAS_EPSILON = np.float32(1e-7)


def topk_indices(values: np.ndarray, k: int) -> np.ndarray:
    """Indices of the k largest entries, lower index first among equal values (tf.nn.top_k)."""
    order = np.argsort(-values, kind="stable")
    return order[:k]


# ------------------------------------------------------------------------------ torch (autograd)

def loss_torch(y_true, y_pred, classification: bool):
    """y_true (N,h,w,1) integer array; y_pred torch tensor (N,h,w,1+C), requires_grad allowed.
    Returns (loss, dict of detached parts).  losses.py:33-126."""
    import torch
    dt = y_pred.dtype
    yt = torch.as_tensor(np.asarray(y_true)).to(torch.int64)
    t = (yt > 0).to(dt)                                                   # losses.py:28
    p = torch.sigmoid(y_pred[..., :1])                                    # losses.py:29
    eps = torch.tensor(1e-7, dtype=dt)
    one = torch.tensor(1.0, dtype=dt)
    pc = torch.clamp(p, eps, one - eps)                                   # K.binary_crossentropy
    x = torch.log(pc / (1 - pc))
    # tf.nn.sigmoid_cross_entropy_with_logits builds relu / -|x| with tf.where(x >= 0, ...), so its
    # gradient at exactly x = 0 is the smooth sigmoid(x) - t (torch.relu / torch.abs would give a
    # different sub-gradient there)
    zeros = torch.zeros_like(x)
    relu_x = torch.where(x >= 0, x, zeros)
    neg_abs_x = torch.where(x >= 0, -x, x)
    ce = relu_x - x * t + torch.log1p(torch.exp(neg_abs_x))                # losses.py:97
    npos = torch.clamp(t.sum(), min=1.0)                                  # losses.py:99
    pos = (ce * t).sum() / npos                                           # losses.py:101
    neg_mask = 1 - t
    ce_neg = ce * neg_mask                                                # losses.py:104
    nneg = torch.clamp(neg_mask.sum(), min=1.0)                           # losses.py:105
    neg = ce_neg.sum() / nneg                                             # losses.py:107
    k = int(min(float(npos), float(nneg)))                                # losses.py:110-112
    flat = ce_neg.reshape(-1)
    idx = torch.as_tensor(topk_indices(flat.detach().numpy(), k).copy())
    hard = flat[idx].mean()                                               # losses.py:116
    if torch.isnan(hard):                                                 # losses.py:117-121
        hard = torch.zeros((), dtype=dt)
    det = L_POSITIVE_WEIGHT * pos + L_NEGATIVE_WEIGHT * neg + L_HARD_NEGATIVE_WEIGHT * hard
    parts = dict(positive=float(pos.detach()), negative=float(neg.detach()),
                 hard_negative=float(hard.detach()), detection=float(det.detach()),
                 n_pos=float(t.sum()), n_neg=float(neg_mask.sum()), k=k)
    if not classification:
        return det, parts
    m = (yt > 0)                                                          # losses.py:76
    lab = ((yt - 1) * m.to(torch.int64)).squeeze(-1)                      # losses.py:78-80
    logits = y_pred[..., 1:]
    cls_ce = torch.nn.functional.cross_entropy(
        logits.reshape(-1, logits.shape[-1]), lab.reshape(-1), reduction="none")
    mf = m.to(dt).reshape(-1)
    cls = (cls_ce * mf).sum() / torch.clamp(mf.sum(), min=1.0)            # losses.py:82
    parts["classification"] = float(cls.detach())
    loss = L_DETECTION_WEIGHT * det + L_CLASSIFICATION_WEIGHT * cls        # losses.py:60
    return loss, parts


def loss_and_grad(y_true, y_pred, classification: bool, dtype="float32"):
    """NumPy in/out convenience: (loss float, parts, dL/dy_pred ndarray)."""
    import torch
    dt = getattr(torch, dtype)
    yp = torch.tensor(np.asarray(y_pred), dtype=dt, requires_grad=True)
    loss, parts = loss_torch(y_true, yp, classification)
    loss.backward()
    return float(loss.detach()), parts, yp.grad.numpy()


# ------------------------------------------------------------------------------ NumPy, hand-derived

def loss_and_grad_numpy(y_true, y_pred, classification: bool):
    """Independent hand-derived restatement (float64 accumulation) used to cross-check autograd.

    d ce/d z = (sigmoid(x) - t) * [eps <= sigmoid(z) <= 1-eps]  (x = re-logit of the clipped p;
    the chain log(p/(1-p)) o sigmoid has unit derivative)."""
    yp = np.asarray(y_pred, dtype=np.float32)
    yt = np.asarray(y_true).astype(np.int64)
    z = yp[..., :1]
    t = (yt > 0).astype(np.float32)
    p = (1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    lo, hi = KERAS_EPSILON, np.float32(1.0) - KERAS_EPSILON
    pc = np.clip(p, lo, hi)
    x = np.log(pc / (np.float32(1) - pc))
    ce = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
    open_ = ((p >= lo) & (p <= hi)).astype(np.float64)
    dce = (1.0 / (1.0 + np.exp(-x.astype(np.float64))) - t) * open_
    npos = max(float(t.sum(dtype=np.float64)), 1.0)
    nneg = max(float((1 - t).sum(dtype=np.float64)), 1.0)
    ce64 = ce.astype(np.float64)
    pos = (ce64 * t).sum() / npos
    ce_neg = ce * (1 - t)
    neg = ce_neg.astype(np.float64).sum() / nneg
    k = int(min(npos, nneg))
    idx = topk_indices(ce_neg.reshape(-1), k)
    hard = ce_neg.reshape(-1)[idx].astype(np.float64).mean()
    sel = np.zeros(ce_neg.size, dtype=np.float64)
    sel[idx] = 1.0 / k
    wgt = (L_POSITIVE_WEIGHT * t / npos + L_NEGATIVE_WEIGHT * (1 - t) / nneg
           + L_HARD_NEGATIVE_WEIGHT * sel.reshape(t.shape) * (1 - t))
    grad = np.zeros(yp.shape, dtype=np.float64)
    grad[..., :1] = wgt * dce
    det = L_POSITIVE_WEIGHT * pos + L_NEGATIVE_WEIGHT * neg + L_HARD_NEGATIVE_WEIGHT * hard
    parts = dict(positive=pos, negative=neg, hard_negative=hard, detection=det, k=k)
    loss = det
    if classification:
        m = (yt > 0)[..., 0]
        lab = ((yt - 1) * (yt > 0))[..., 0]
        lg = yp[..., 1:].astype(np.float64)
        lg = lg - lg.max(-1, keepdims=True)
        lse = np.log(np.exp(lg).sum(-1))
        ce_c = lse - np.take_along_axis(lg, lab[..., None], -1)[..., 0]
        nm = max(float(m.sum()), 1.0)
        cls = (ce_c * m).sum() / nm
        sm = np.exp(lg - lse[..., None])
        onehot = np.zeros_like(sm)
        np.put_along_axis(onehot, lab[..., None], 1.0, -1)
        grad[..., 1:] = (sm - onehot) * (m[..., None] / nm) * L_CLASSIFICATION_WEIGHT
        parts["classification"] = cls
        loss = L_DETECTION_WEIGHT * det + L_CLASSIFICATION_WEIGHT * cls
        grad[..., :1] *= L_DETECTION_WEIGHT
    return float(loss), parts, grad


# ------------------------------------------------------------------------------ optimizer

def adam_step(params, grads, m, v, t, lr=1e-3, beta_1=0.9, beta_2=0.999, eps=1e-7):
    """Keras-2 Adam (train.py:110), in-place on float32 arrays; t is the 1-based step count."""
    lr_t = np.float32(lr * np.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t))
    for p, g, mi, vi in zip(params, grads, m, v):
        g = np.asarray(g, dtype=np.float32)
        mi[...] = np.float32(beta_1) * mi + np.float32(1 - beta_1) * g
        vi[...] = np.float32(beta_2) * vi + np.float32(1 - beta_2) * g * g
        p[...] = p - lr_t * mi / (np.sqrt(vi) + np.float32(eps))


def train_step_torch(weights, images, y_true, classification, fml_compatible=True, dtype="float32"):
    """One forward+backward of the whole graph (gradient oracle for the CUDA backward):
    returns (loss, parts, grads in get_weights() order / shapes)."""
    import torch
    from . import net
    dt = getattr(torch, dtype)
    p = net.torch_params(weights, dt)
    for v_ in p.values():
        v_.requires_grad_(True)
    x = torch.as_tensor(np.ascontiguousarray(images)).to(dt).permute(0, 3, 1, 2).contiguous()
    y = net.forward_torch_nchw(p, x, fml_compatible).permute(0, 2, 3, 1)
    loss, parts = loss_torch(y_true, y, classification)
    loss.backward()
    g = []
    for i in range(3):
        g.append(p[f"dw{i}"].grad.permute(2, 3, 0, 1).numpy())      # (3,3,C,1)
        g.append(p[f"pw{i}"].grad.permute(2, 3, 1, 0).numpy())      # (1,1,C,F)
        g.append(p[f"b{i}"].grad.numpy())
    for li in range(6):
        g.append(p[f"k{li}"].grad.permute(2, 3, 1, 0).numpy())      # HWIO
        g.append(p[f"kb{li}"].grad.numpy())
    g.append(p["hk"].grad.permute(2, 3, 1, 0).numpy())
    g.append(p["hb"].grad.numpy())
    return float(loss.detach()), parts, [np.ascontiguousarray(a) for a in g], y.detach().numpy()
