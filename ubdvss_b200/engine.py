"""Thin Python owner of one ``ubd_handle`` (one GPU, one stream).  Everything numeric happens in
libubd.so; this file only marshals NumPy buffers across the C ABI (include/ubd.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import COMPONENT_DTYPE, check, ptr


def _weight_sizes(grey: bool, n_classes: int):
    cin = 1 if grey else 3
    sizes = []
    for c in (cin, 24, 24):
        sizes += [9 * c, c * 24, 24]
    for _ in range(6):
        sizes += [9 * 24 * 24, 24]
    sizes += [24 * (1 + n_classes), 1 + n_classes]
    return sizes


def weight_shapes(grey: bool = True, n_classes: int = 0):
    """Keras ``get_weights()`` shapes of net.py:286-313."""
    cin = 1 if grey else 3
    shapes = []
    for c in (cin, 24, 24):
        shapes += [(3, 3, c, 1), (1, 1, c, 24), (24,)]
    for _ in range(6):
        shapes += [(3, 3, 24, 24), (24,)]
    shapes += [(1, 1, 24, 1 + n_classes), (1 + n_classes,)]
    return shapes


class Engine:
    def __init__(self, device: int = 0, grey: bool = True, fml_compatible: bool = True, n_classes: int = 0,
                 precision: str = "fp32"):
        self._lib = _lib.load()
        self.grey, self.fml_compatible, self.n_classes = bool(grey), bool(fml_compatible), int(n_classes)
        self.cin = 1 if grey else 3
        self.n_out = 1 + self.n_classes
        self.precision = precision
        h = C.c_void_p()
        rc = self._lib.ubd_create(int(device), int(grey), int(fml_compatible), int(n_classes),
                                  _lib.PRECISIONS[precision], C.byref(h))
        if rc != 0:
            raise _lib.UbdError(rc, self._lib.ubd_last_error(None).decode())
        self._h = h
        self._sizes = _weight_sizes(grey, n_classes)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ubd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- weights
    def _weight_ptrs(self, arrays):
        n = _lib.N_WEIGHT_ARRAYS
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
        sizes = (C.c_int64 * n)(*[a.size for a in arrays])
        return ptrs, sizes

    def set_weights(self, weights):
        if len(weights) != _lib.N_WEIGHT_ARRAYS:
            raise ValueError(f"expected {_lib.N_WEIGHT_ARRAYS} weight arrays (Keras get_weights() order), got {len(weights)}")
        arrays = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        for a, s in zip(arrays, weight_shapes(self.grey, self.n_classes)):
            if tuple(a.shape) != tuple(s):
                raise ValueError(f"weight shape {a.shape} does not match the layer's {s}")
        ptrs, sizes = self._weight_ptrs(arrays)
        check(self._h, self._lib.ubd_set_weights(self._h, ptrs, sizes, len(arrays)))

    def get_weights(self):
        arrays = [np.empty(s, np.float32) for s in weight_shapes(self.grey, self.n_classes)]
        ptrs, sizes = self._weight_ptrs(arrays)
        check(self._h, self._lib.ubd_get_weights(self._h, ptrs, sizes, len(arrays)))
        return arrays

    def get_grads(self):
        arrays = [np.empty(s, np.float32) for s in weight_shapes(self.grey, self.n_classes)]
        ptrs, sizes = self._weight_ptrs(arrays)
        check(self._h, self._lib.ubd_get_grads(self._h, ptrs, sizes, len(arrays)))
        return arrays

    def set_option(self, name: str, value: int):
        check(self._h, self._lib.ubd_set_option(self._h, name.encode(), int(value)))

    # ---------------------------------------------------------------- inference
    def _images(self, images):
        x = np.asarray(images)
        if x.ndim != 4 or x.shape[3] != self.cin:
            raise ValueError(f"images must be (N,H,W,{self.cin}), got {x.shape}")
        if x.dtype == np.uint8:
            return np.ascontiguousarray(x), _lib.UBD_U8
        return np.ascontiguousarray(x, dtype=np.float32), _lib.UBD_F32

    def forward(self, images, preproc: int = _lib.PREPROC_NONE):
        x, dt = self._images(images)
        n, H, W, _ = x.shape
        out = np.empty((n, H // 4, W // 4, self.n_out), np.float32)
        check(self._h, self._lib.ubd_forward(self._h, ptr(x), dt, n, H, W, preproc, ptr(out)))
        return out

    def segment(self, images, logit_thr: float, min_area_x2: int, preproc: int = _lib.PREPROC_NONE,
                want_logits: bool = True, want_labels: bool = False, max_comps: int = 0):
        """-> (mask uint8 (N,h,w), logits | None, labels | None, comps structured array, counts int32[N])."""
        x, dt = self._images(images)
        n, H, W, _ = x.shape
        h4, w4 = H // 4, W // 4
        mask = np.empty((n, h4, w4), np.uint8)
        logits = np.empty((n, h4, w4, self.n_out), np.float32) if want_logits else None
        labels = np.empty((n, h4, w4), np.int32) if want_labels else None
        cap = max_comps or max(1024, 64 * n)
        while True:
            comps = np.zeros(cap, COMPONENT_DTYPE)
            counts = np.zeros(n, np.int32)
            rc = self._lib.ubd_segment(self._h, ptr(x), dt, n, H, W, preproc, np.float32(logit_thr), int(min_area_x2),
                                       ptr(mask), ptr(logits), ptr(labels), ptr(comps), cap, ptr(counts))
            if rc == -4 and not max_comps and cap < (1 << 22) and b"capacity" in self._lib.ubd_last_error(self._h):
                cap *= 8
                continue
            check(self._h, rc)
            break
        return mask, logits, labels, comps[:int(counts.sum())], counts

    def postprocess(self, mask, cls_logits=None, min_area_x2: int = 10, want_labels: bool = False, max_comps: int = 0):
        """mask (N,h,w) -> (labels | None, comps, counts).  SegmapManager.postprocess on the GPU."""
        m = np.ascontiguousarray(mask, dtype=np.uint8)
        if m.ndim != 3:
            raise ValueError("mask must be (N,h,w)")
        n, mh, mw = m.shape
        n_cls = 0
        cl = None
        if cls_logits is not None and np.shape(cls_logits)[-1] > 0:
            cl = np.ascontiguousarray(cls_logits, dtype=np.float32)
            n_cls = cl.shape[-1]
            if cl.shape != (n, mh, mw, n_cls):
                raise ValueError("class logits must be (N,h,w,C)")
        labels = np.empty((n, mh, mw), np.int32) if want_labels else None
        cap = max_comps or max(1024, 64 * n)
        while True:
            comps = np.zeros(cap, COMPONENT_DTYPE)
            counts = np.zeros(n, np.int32)
            rc = self._lib.ubd_postprocess(self._h, ptr(m), ptr(cl), n, mh, mw, n_cls, int(min_area_x2),
                                           ptr(labels), ptr(comps), cap, ptr(counts))
            if rc == -4 and not max_comps and cap < (1 << 22) and b"capacity" in self._lib.ubd_last_error(self._h):
                cap *= 8
                continue
            check(self._h, rc)
            break
        return labels, comps[:int(counts.sum())], counts

    # ---------------------------------------------------------------- training
    def train_step(self, images, y_true, preproc: int = _lib.PREPROC_NONE):
        x, dt = self._images(images)
        n, H, W, _ = x.shape
        y = np.ascontiguousarray(np.asarray(y_true).reshape(n, H // 4, W // 4), dtype=np.int32)
        parts = np.zeros(6, np.float32)
        check(self._h, self._lib.ubd_train_step(self._h, ptr(x), dt, n, H, W, preproc, ptr(y), ptr(parts)))
        return parts

    def train_update(self, images, y_true, preproc: int = _lib.PREPROC_NONE, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        """train_step + gradient all-reduce (if ``comm_init`` was called) + Adam with one host synchronisation."""
        x, dt = self._images(images)
        n, H, W, _ = x.shape
        y = np.ascontiguousarray(np.asarray(y_true).reshape(n, H // 4, W // 4), dtype=np.int32)
        parts = np.zeros(6, np.float32)
        check(self._h, self._lib.ubd_train_update(self._h, ptr(x), dt, n, H, W, preproc, ptr(y), lr, beta_1, beta_2, epsilon, ptr(parts)))
        return parts

    def loss(self, logits, y_true):
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        n, mh, mw, c = lg.shape
        if c != self.n_out:
            raise ValueError("logits channel count does not match the model head")
        y = np.ascontiguousarray(np.asarray(y_true).reshape(n, mh, mw), dtype=np.int32)
        parts = np.zeros(6, np.float32)
        dl = np.empty_like(lg)
        check(self._h, self._lib.ubd_loss(self._h, ptr(lg), ptr(y), n, mh, mw, ptr(parts), ptr(dl)))
        return parts, dl

    def grad_buffer(self):
        p = C.c_void_p()
        n = C.c_int64()
        check(self._h, self._lib.ubd_grad_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---------------------------------------------------------------- data-parallel exchange (NCCL inside the library)
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id; create on rank 0 and hand to every rank (any channel)."""
        buf = np.zeros(128, np.uint8)
        check(None, _lib.load().ubd_comm_unique_id(ptr(buf)))
        return buf.tobytes()

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = np.frombuffer(unique_id, np.uint8).copy()
        assert buf.size == 128
        check(self._h, self._lib.ubd_comm_init(self._h, ptr(buf), int(rank), int(world)))
        self.world = int(world)

    def comm_destroy(self):
        check(self._h, self._lib.ubd_comm_destroy(self._h))
        self.world = 1

    def allreduce_grads(self):
        """Sum of the flat gradient buffer over the ranks, in place, on the handle's stream."""
        check(self._h, self._lib.ubd_allreduce_grads(self._h))

    def metric_counts(self):
        """(tp, tn, fp, fn, cls_correct, cls_total) of the last train_step / loss batch (keras_metrics.py:116-191)."""
        c = np.zeros(6, np.int64)
        check(self._h, self._lib.ubd_metric_counts(self._h, ptr(c)))
        return c

    def adam_step(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, grad_scale=1.0):
        check(self._h, self._lib.ubd_adam_step(self._h, lr, beta_1, beta_2, epsilon, grad_scale))

    def synchronize(self):
        check(self._h, self._lib.ubd_synchronize(self._h))

    def debug_dilated_layer(self, x_nhwc, layer: int, precision: str = "fp32"):
        x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
        n, mh, mw, c = x.shape
        assert c == 24
        out = np.empty_like(x)
        check(self._h, self._lib.ubd_debug_dilated_layer(self._h, ptr(x), ptr(out), layer, n, mh, mw, _lib.PRECISIONS[precision]))
        return out

    def debug_wgrad(self, x_nhwc, g_nhwc, dilation: int):
        """(dK (3,3,24,24), dB (24,)) of one dilated layer from its input map and output gradient (tcgen05 wgrad kernel)."""
        x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
        g = np.ascontiguousarray(g_nhwc, dtype=np.float32)
        n, mh, mw, c = x.shape
        assert c == 24 and g.shape == x.shape
        dk = np.empty((3, 3, 24, 24), np.float32)
        db = np.empty(24, np.float32)
        check(self._h, self._lib.ubd_debug_wgrad(self._h, ptr(x), ptr(g), n, mh, mw, int(dilation), ptr(dk), ptr(db)))
        return dk, db

    def set_stream(self, cuda_stream: int | None):
        check(self._h, self._lib.ubd_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def stat(self, name: str) -> float:
        v = C.c_double()
        check(self._h, self._lib.ubd_get_stat(self._h, name.encode(), C.byref(v)))
        return v.value

    def segment_dev(self, d_images_ptr: int, dtype: int, n: int, H: int, W: int, logit_thr: float, min_area_x2: int,
                    preproc: int = _lib.PREPROC_NONE, d_mask_ptr: int = 0, d_logits_ptr: int = 0, max_comps: int = 0):
        """Device-resident input (raw pointers, e.g. torch ``data_ptr()``) -> (comps, counts)."""
        cap = max_comps or max(1024, 64 * n)
        comps = np.zeros(cap, COMPONENT_DTYPE)
        counts = np.zeros(n, np.int32)
        check(self._h, self._lib.ubd_segment_dev(self._h, C.c_void_p(d_images_ptr), dtype, n, H, W, preproc,
                                                 np.float32(logit_thr), int(min_area_x2), C.c_void_p(d_mask_ptr or 0),
                                                 C.c_void_p(d_logits_ptr or 0), ptr(comps), cap, ptr(counts)))
        return comps[:int(counts.sum())], counts

    # ---------------------------------------------------------------- input side
    def prepare_images(self, images, out_h: int, out_w: int, to_grey: bool = True):
        """(N,H,W,1|3) uint8 -> (N,out_h,out_w,1|3) uint8: ``Image.resize((out_w, out_h), Image.BICUBIC)`` then
        ``convert('L')`` (if ``to_grey`` and the input is RGB), bit-identical to Pillow, on the GPU."""
        x = np.ascontiguousarray(images, dtype=np.uint8)
        if x.ndim != 4 or x.shape[3] not in (1, 3):
            raise ValueError(f"images must be (N,H,W,1|3) uint8, got {x.shape}")
        n, H, W, c = x.shape
        out = np.empty((n, out_h, out_w, 1 if (to_grey and c == 3) else c), np.uint8)
        check(self._h, self._lib.ubd_prepare_images(self._h, ptr(x), n, H, W, c, int(out_h), int(out_w), int(bool(to_grey)), ptr(out)))
        return out

    # ---------------------------------------------------------------- pipelined inference (up to three batches in flight)
    def segment_submit(self, images, logit_thr: float, min_area_x2: int, preproc: int = _lib.PREPROC_NONE,
                       mask_out=None, logits_out=None, max_comps: int = 0, device_ptr: int = 0, shape=None, dtype=None):
        """Queue one batch and return a ticket object for ``segment_wait``.  ``images``: host array (pass pinned memory
        for an asynchronous copy), or ``device_ptr`` + ``shape`` (n, H, W) + ``dtype`` for device-resident input.
        ``mask_out`` / ``logits_out``: optional host arrays filled by the time ``segment_wait`` returns."""
        if device_ptr:
            n, H, W = shape
            dt = dtype
            x = None
        else:
            x, dt = self._images(images)
            n, H, W, _ = x.shape
        cap = max_comps or max(1024, 64 * n)
        t = C.c_int()
        if device_ptr:
            check(self._h, self._lib.ubd_segment_submit_dev(self._h, C.c_void_p(device_ptr), dt, n, H, W, preproc, np.float32(logit_thr),
                                                            int(min_area_x2), cap, C.byref(t)))
        else:
            check(self._h, self._lib.ubd_segment_submit(self._h, ptr(x), dt, n, H, W, preproc, np.float32(logit_thr), int(min_area_x2),
                                                        ptr(mask_out), ptr(logits_out), cap, C.byref(t)))
        # the ticket keeps the host buffers alive until the batch has been collected
        return {"ticket": t.value, "n": n, "cap": cap, "keep": (x, mask_out, logits_out)}

    def segment_wait(self, ticket):
        """-> (comps, counts) of a submitted batch (tickets are collected in submission order)."""
        comps = np.zeros(ticket["cap"], COMPONENT_DTYPE)
        counts = np.zeros(ticket["n"], np.int32)
        check(self._h, self._lib.ubd_segment_wait(self._h, ticket["ticket"], ptr(comps), ticket["cap"], ptr(counts)))
        ticket["keep"] = None
        return comps[:int(counts.sum())], counts

    def launch_count(self) -> int:
        return int(self._lib.ubd_launch_count(self._h))

    @property
    def handle(self):
        return self._h


def min_area_box(points_xy) -> np.ndarray:
    """cv2.boxPoints(cv2.minAreaRect(points)) on the host (no GPU needed)."""
    pts = np.ascontiguousarray(np.asarray(points_xy).reshape(-1, 2), dtype=np.int32)
    box = np.zeros(8, np.float32)
    rc = _lib.load().ubd_min_area_box(ptr(pts), pts.shape[0], ptr(box))
    if rc != 0:
        raise _lib.UbdError(rc, "ubd_min_area_box")
    return box
