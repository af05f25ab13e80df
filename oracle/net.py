"""CPU restatement of the reference's dilated FCN forward pass.  TEST INFRASTRUCTURE ONLY.

Follows ``semantic_segmentation/net.py:225-252`` (``conv_bn``) and ``net.py:278-314``
(``_build_dilated_conv_model``).  PARITY UNPINNED: the reference delegates the arithmetic
to Keras-2.x / TensorFlow-1.x (absent here, ``requirements.txt:2,6``); the semantics
encoded below are the documented Keras ones:

* ``Conv2D`` is a cross-correlation (no kernel flip), kernel layout HWIO, bias on.
* ``SeparableConv2D`` = depthwise 3x3 (kernel ``(3,3,Cin,1)``, depth_multiplier 1, no bias,
  no activation) followed by pointwise 1x1 ``(1,1,Cin,F)`` + bias + activation.
* ``padding='same'`` with stride 1 and dilation d pads d zeros on every side.
* ``use_strides_compatible_with_fml`` (``net.py:229-232``): ZeroPadding2D(((1,0),(1,0)))
  then stride-2 ``'valid'``; without it TF ``'same'`` stride 2 pads bottom/right for even sizes.
* weights travel as the list ``model.get_weights()`` returns (``net.py:418-427``), 23 arrays.

Two independent implementations are provided (NumPy shifted slices, torch-CPU conv2d);
``tests/test_oracle_net.py`` checks they agree and checks both against analytic cases.
"""
from __future__ import annotations

import numpy as np

N_FILTERS = 24                      # net.py:289
DILATIONS = (1, 2, 4, 8, 16, 1)     # net.py:298-304
SCALE = 4                           # net.py:314


def weight_spec(n_classes: int = 0, grey: bool = True):
    """(keras_name, shape) for the 23 arrays in ``get_weights()`` order."""
    cin = 1 if grey else 3          # net.py:286
    spec = []
    for i, c in enumerate((cin, N_FILTERS, N_FILTERS), start=1):   # net.py:292-296
        spec.append((f"separable_conv2d_{i}/depthwise_kernel", (3, 3, c, 1)))
        spec.append((f"separable_conv2d_{i}/pointwise_kernel", (1, 1, c, N_FILTERS)))
        spec.append((f"separable_conv2d_{i}/bias", (N_FILTERS,)))
    for i in range(1, 7):                                           # net.py:298-304
        spec.append((f"conv2d_{i}/kernel", (3, 3, N_FILTERS, N_FILTERS)))
        spec.append((f"conv2d_{i}/bias", (N_FILTERS,)))
    spec.append(("conv2d_7/kernel", (1, 1, N_FILTERS, 1 + n_classes)))   # net.py:311
    spec.append(("conv2d_7/bias", (1 + n_classes,)))
    return spec


def init_weights(n_classes: int = 0, seed: int = 1234, grey: bool = True, bias_std: float = 0.1):
    """Glorot-uniform kernels (Keras default, net.py:226); biases N(0, bias_std) so the bias
    path is exercised (Keras would start them at zero)."""
    rng = np.random.default_rng(seed)
    out = []
    for name, shape in weight_spec(n_classes, grey):
        if len(shape) == 1:
            out.append((rng.standard_normal(shape) * bias_std).astype(np.float32))
        else:
            kh, kw, cin, cout = shape
            limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
            out.append(rng.uniform(-limit, limit, size=shape).astype(np.float32))
    return out


def preprocess(images, kind: str = "none"):
    """net.py:163-169,217-218."""
    if kind == "none":
        return images
    if kind == "mobilenet_like":
        return (images - 127.5) / 127.5
    raise ValueError("Unknown preprocessing type")


# --------------------------------------------------------------------------------------
# implementation 1: NumPy shifted slices
# --------------------------------------------------------------------------------------

def _pad(x, top, bottom, left, right):
    return np.pad(x, ((0, 0), (top, bottom), (left, right), (0, 0)))


def _depthwise3x3_np(x, k, stride, pads):
    """x (N,H,W,C), k (3,3,C,1); pads=(top,bottom,left,right); 'valid' after padding."""
    xp = _pad(x, *pads)
    hp, wp = xp.shape[1:3]
    ho = (hp - 3) // stride + 1
    wo = (wp - 3) // stride + 1
    out = np.zeros((x.shape[0], ho, wo, x.shape[3]), dtype=x.dtype)
    for i in range(3):
        for j in range(3):
            sl = xp[:, i:i + stride * (ho - 1) + 1:stride, j:j + stride * (wo - 1) + 1:stride, :]
            out += sl * k[i, j, :, 0]
    return out


def _conv3x3_np(x, k, d):
    """'same' 3x3 conv, dilation d: out[y,x,o] = sum in[y+(i-1)d, x+(j-1)d, c] k[i,j,c,o]."""
    n, h, w, _ = x.shape
    xp = _pad(x, d, d, d, d)
    out = np.zeros((n, h, w, k.shape[3]), dtype=x.dtype)
    for i in range(3):
        for j in range(3):
            out += xp[:, i * d:i * d + h, j * d:j * d + w, :] @ k[i, j]
    return out


def _stride2_pads(h, w, fml_compatible):
    if fml_compatible:                       # net.py:229-232: top/left zero row/col, then 'valid'
        return (1, 0, 1, 0)
    # TF 'same', stride 2, kernel 3: total pad = max((ceil(H/2)-1)*2 + 3 - H, 0); top = total//2
    def one(sz):
        total = max((-(-sz // 2) - 1) * 2 + 3 - sz, 0)
        return total // 2, total - total // 2
    t, b = one(h)
    l, r = one(w)
    return (t, b, l, r)


def forward_numpy(weights, images, fml_compatible=True, dtype=np.float32, return_all=False):
    """images (N,H,W,Cin) already preprocessed -> logits (N,H/4,W/4,1+C).  net.py:286-313."""
    w = [np.asarray(a, dtype=dtype) for a in weights]
    x = np.asarray(images, dtype=dtype)
    acts = []
    relu = lambda a: np.maximum(a, 0)
    # L1: separable s2
    x = _depthwise3x3_np(x, w[0], 2, _stride2_pads(x.shape[1], x.shape[2], fml_compatible))
    x = relu(x @ w[1][0, 0] + w[2]); acts.append(x)
    # L2: separable s1 'same'
    x = _depthwise3x3_np(x, w[3], 1, (1, 1, 1, 1))
    x = relu(x @ w[4][0, 0] + w[5]); acts.append(x)
    # L3: separable s2
    x = _depthwise3x3_np(x, w[6], 2, _stride2_pads(x.shape[1], x.shape[2], fml_compatible))
    x = relu(x @ w[7][0, 0] + w[8]); acts.append(x)
    # L4..L9
    for li, d in enumerate(DILATIONS):
        x = relu(_conv3x3_np(x, w[9 + 2 * li], d) + w[10 + 2 * li]); acts.append(x)
    # L10 head, linear
    x = x @ w[21][0, 0] + w[22]; acts.append(x)
    return (x, acts) if return_all else x


# --------------------------------------------------------------------------------------
# implementation 2: torch-CPU conv2d (also the timed CPU baseline in bench.py)
# --------------------------------------------------------------------------------------

def torch_params(weights, dtype=None):
    import torch
    dtype = dtype or torch.float32
    t = [torch.as_tensor(np.asarray(a), dtype=dtype) for a in weights]
    p = {}
    for i in range(3):
        dw, pw, b = t[3 * i:3 * i + 3]
        p[f"dw{i}"] = dw.permute(2, 3, 0, 1).contiguous()          # (C,1,3,3)
        p[f"pw{i}"] = pw.permute(3, 2, 0, 1).contiguous()          # (F,C,1,1)
        p[f"b{i}"] = b
    for li in range(6):
        p[f"k{li}"] = t[9 + 2 * li].permute(3, 2, 0, 1).contiguous()   # OIHW
        p[f"kb{li}"] = t[10 + 2 * li]
    p["hk"] = t[21].permute(3, 2, 0, 1).contiguous()
    p["hb"] = t[22]
    return p


def forward_torch_nchw(p, x, fml_compatible=True, return_all=False):
    """x: (N,Cin,H,W) tensor -> logits (N,1+C,H/4,W/4).  Differentiable (gradient oracle)."""
    import torch.nn.functional as F
    acts = []

    def sep(x, i, stride):
        c = x.shape[1]
        if stride == 2:
            t, b, l, r = _stride2_pads(x.shape[2], x.shape[3], fml_compatible)
        else:
            t = b = l = r = 1
        x = F.pad(x, (l, r, t, b))
        x = F.conv2d(x, p[f"dw{i}"], None, stride=stride, groups=c)
        x = F.conv2d(x, p[f"pw{i}"], p[f"b{i}"])
        return F.relu(x)

    x = sep(x, 0, 2); acts.append(x)
    x = sep(x, 1, 1); acts.append(x)
    x = sep(x, 2, 2); acts.append(x)
    for li, d in enumerate(DILATIONS):
        x = F.relu(F.conv2d(x, p[f"k{li}"], p[f"kb{li}"], padding=d, dilation=d)); acts.append(x)
    x = F.conv2d(x, p["hk"], p["hb"]); acts.append(x)
    return (x, acts) if return_all else x


def forward_torch(weights, images, fml_compatible=True, dtype=None):
    """NHWC numpy in, NHWC numpy logits out."""
    import torch
    dtype = dtype or torch.float32
    with torch.no_grad():
        p = torch_params(weights, dtype)
        x = torch.as_tensor(np.ascontiguousarray(images)).to(dtype).permute(0, 3, 1, 2).contiguous()
        y = forward_torch_nchw(p, x, fml_compatible)
        return y.permute(0, 2, 3, 1).contiguous().numpy()


def round_tf32(a, mode="rna"):
    """Round fp32 to the 10-bit-mantissa TF32 grid (used to budget the tensor-core path)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    if mode == "rz":
        return (u & np.uint32(0xFFFFE000)).view(np.float32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
