"""Model-side boundary of the reference (semantic_segmentation/net.py): ``NetConfig``,
``PreprocessingType``, ``NetManager`` and a Keras-``Model``-shaped object whose arithmetic runs in
libubd.so.  Same names, argument meaning and error behaviour as the reference for everything the
hot path touches; the alternative architectures of net.py:316-413 are
never selected by ``build_model`` (net.py:273-276) and are not provided.
"""
from __future__ import annotations

import copy
import io
import logging
import os
import pickle
from enum import Enum

import numpy as np

from . import _lib
from .engine import Engine, weight_shapes


class PreprocessingType(Enum):          # net.py:62-64
    NONE = 0
    MOBILENET_LIKE = 1


supported_preprocessing_types = {       # net.py:67-70
    "none": PreprocessingType.NONE,
    "mobilenet_like": PreprocessingType.MOBILENET_LIKE,
}

_PREPROC_CODE = {PreprocessingType.NONE: _lib.PREPROC_NONE, PreprocessingType.MOBILENET_LIKE: _lib.PREPROC_MOBILENET,
                 "none": _lib.PREPROC_NONE, "mobilenet_like": _lib.PREPROC_MOBILENET, None: _lib.PREPROC_NONE}


def preprocess_image_mobilenet(image):      # net.py:217-218
    return (image - 127.5) / 127.5


def depreprocess_image_mobilenet(image):    # net.py:221-222
    return image * 127.5 + 127.5


class NetConfig:
    """Configuration of the net and the pipeline around it (net.py:73-214).  Attribute names match
    the reference so that its pickled ``config.pkl`` (net.py:468-472) restores into this class."""

    @staticmethod
    def from_others(base_config, side_multiple=None, max_image_side=None, min_pixels_for_detection=None):
        new_config = copy.deepcopy(base_config)
        if side_multiple:
            new_config._side_multiple = side_multiple
        if max_image_side:
            new_config._max_side = max_image_side
        if min_pixels_for_detection:
            new_config._min_pixels_for_detection = min_pixels_for_detection
        return new_config

    def __init__(self, object_types_fname=None, scale=4, fml_compatible=True, no_classification=False,
                 side_multiple=64, max_image_side=512, min_pixels_for_detection=5,
                 preprocessing=PreprocessingType.NONE, grey=True):
        if object_types_fname is None:
            self._class_names = None
            self._is_classification_supported = False
        else:
            self._is_classification_supported = not no_classification
            self._read_classnames_from_file(object_types_fname)
        self._grey = grey
        self._scale = scale
        self._fml_compatible = fml_compatible
        self._preprocessing = preprocessing
        self._side_multiple = side_multiple
        self._max_side = max_image_side
        self._min_pixels_for_detection = min_pixels_for_detection

    def log_classification_mode(self):
        """Called by train.py:97 before training; the wording is this package's own."""
        names = self._class_names
        if self.is_classification_supported():
            logging.info("classification head enabled for %d object types: %s", len(names), names)
        elif names is not None:
            logging.info("detection only (class head disabled); object types of the dataset: %s", names)
        else:
            logging.info("detection only: any barcode of the datasets is a positive")

    def is_grey(self): return self._grey
    def get_scale(self): return self._scale
    def get_min_pixels_for_detection(self): return self._min_pixels_for_detection
    def get_side_multiple(self): return self._side_multiple
    def get_max_side(self): return self._max_side
    def is_fml_compatible(self): return self._fml_compatible
    def get_preprocessing_type(self): return self._preprocessing

    def get_preprocessing_fn(self):
        if self._preprocessing == PreprocessingType.NONE:
            return lambda x: x
        elif self._preprocessing == PreprocessingType.MOBILENET_LIKE:
            return preprocess_image_mobilenet
        raise ValueError("Unknown preprocessing type")

    def get_depreprocessing_fn(self):
        if self._preprocessing == PreprocessingType.NONE:
            return lambda x: x
        elif self._preprocessing == PreprocessingType.MOBILENET_LIKE:
            return depreprocess_image_mobilenet
        raise ValueError("Unknown preprocessing type")

    def get_class_names(self): return self._class_names
    def get_n_classes(self): return len(self._class_names)
    def get_class_name(self, class_id): return self._class_names[class_id]
    def get_class_id(self, class_name): return self._class_name_to_id[class_name]
    def is_class_supported(self, class_name): return self._class_names is None or class_name in self._class_name_to_id
    def is_classification_supported(self): return self._is_classification_supported

    def set_class_names(self, class_names, no_classification=False):
        """Convenience not in the reference: class names without going through a file."""
        self._class_names = list(class_names)
        self._class_name_to_id = dict((n, i) for i, n in enumerate(self._class_names))
        self._is_classification_supported = not no_classification

    def _read_classnames_from_file(self, path):
        assert os.path.exists(path), f"File with object class names {path} does not exist"
        logging.info(f"Reading object types from {path}")
        with open(path, "r") as f:
            names = [line.strip() for line in f if line.strip()]
        self._class_names = names
        self._class_name_to_id = dict((n, i) for i, n in enumerate(names))

    def __str__(self):
        fields = {k.lstrip("_"): v for k, v in vars(self).items() if k.startswith("_") and k != "_class_name_to_id"}
        return "NetConfig(" + ", ".join(f"{k}={v!r}" for k, v in sorted(fields.items())) + ")"


class _ConfigUnpickler(pickle.Unpickler):
    """Restores a ``config.pkl`` written by the reference (class path semantic_segmentation.net.*)."""

    def find_class(self, module, name):
        if module in ("semantic_segmentation.net", "ubdvss_b200.net"):
            return {"NetConfig": NetConfig, "PreprocessingType": PreprocessingType}[name]
        return super().find_class(module, name)


class Adam:
    """Keras-2 ``Adam(lr)`` hyper-parameters (train.py:110); the update itself is ``ubd_adam_step``."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=None, decay=0., amsgrad=False):
        if decay or amsgrad:
            raise NotImplementedError("the reference uses plain Adam(lr) (train.py:110)")
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.epsilon = 1e-7 if epsilon is None else float(epsilon)      # K.epsilon()


_KERAS_NAMES = ([f"separable_conv2d_{i}/{p}" for i in (1, 2, 3) for p in ("depthwise_kernel", "pointwise_kernel", "bias")]
                + [f"conv2d_{i}/{p}" for i in range(1, 7) for p in ("kernel", "bias")]
                + ["conv2d_7/kernel", "conv2d_7/bias"])


class B200Model:
    """Duck-types the Keras ``Model`` the reference builds in ``_build_dilated_conv_model``
    (net.py:278-314): ``predict`` / ``compile`` / ``train_on_batch`` / ``fit_generator`` /
    ``get_weights`` / ``set_weights`` / ``save_weights`` / ``load_weights`` / ``save`` / ``summary``."""

    name = "dilated_conv"

    def __init__(self, net_config: NetConfig, device: int = 0, precision: str = "fp32", weights=None, seed=None):
        self._cfg = net_config
        n_classes = net_config.get_n_classes() if net_config.is_classification_supported() else 0   # net.py:307-310
        self._engine = Engine(device=device, grey=net_config.is_grey(), fml_compatible=net_config.is_fml_compatible(),
                              n_classes=n_classes, precision=precision)
        self.n_classes = n_classes
        self.device = device
        self._optimizer = None
        self._classification_loss = False
        self.metrics_names = ["loss"]
        self.stop_training = False
        self._dist = None
        if weights is None:
            weights = self._glorot_uniform(seed)
        self.set_weights(weights)

    # ---- initialisation: Keras defaults, glorot_uniform kernels and zero biases (net.py:226)
    def _glorot_uniform(self, seed):
        rng = np.random.default_rng(seed)
        out = []
        for shape in weight_shapes(self._cfg.is_grey(), self.n_classes):
            if len(shape) == 1:
                out.append(np.zeros(shape, np.float32))
            else:
                kh, kw, cin, cout = shape
                limit = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
                out.append(rng.uniform(-limit, limit, size=shape).astype(np.float32))
        return out

    @property
    def engine(self) -> Engine:
        return self._engine

    # ---- inference
    def predict(self, x, batch_size=None, verbose=0, steps=None, preprocessing=None):
        """``model.predict(images)`` (model_runner.py:119, predict.py:74-76): (N,H,W,Cin) any real
        dtype -> float32 logits (N,H/4,W/4,1+C).  Float input is taken as already preprocessed, as
        Keras would; ``preprocessing`` folds the config's preprocessing for uint8 input."""
        return self._engine.forward(x, _PREPROC_CODE[preprocessing])

    def segment(self, images, logit_thr, min_area_x2, preprocessing=None, want_labels=False):
        mask, logits, labels, comps, counts = self._engine.segment(
            images, logit_thr, min_area_x2, _PREPROC_CODE[preprocessing], want_logits=True, want_labels=want_labels)
        if want_labels:
            return mask, logits, comps, counts, labels
        return mask, logits, comps, counts

    def segment_submit(self, images, logit_thr, min_area_x2, preprocessing=None):
        """Queue one batch (three may be in flight); see ``ModelRunner.predict_stream``."""
        x = np.asarray(images)
        n, H, W = x.shape[:3]
        mask = np.empty((n, H // 4, W // 4), np.uint8)
        logits = np.empty((n, H // 4, W // 4, 1 + self.n_classes), np.float32)
        t = self._engine.segment_submit(x, logit_thr, min_area_x2, _PREPROC_CODE[preprocessing], mask_out=mask, logits_out=logits)
        t["mask"], t["logits"] = mask, logits
        return t

    def segment_wait(self, ticket):
        comps, counts = self._engine.segment_wait(ticket)
        return ticket["mask"], ticket["logits"], comps, counts

    # ---- weights (net.py:418-427)
    def get_weights(self):
        return self._engine.get_weights()

    def set_weights(self, weights):
        self._engine.set_weights(weights)

    def count_params(self):
        return int(sum(int(np.prod(s)) for s in weight_shapes(self._cfg.is_grey(), self.n_classes)))

    def _keras_layers(self, weights):
        """``model.layers`` of net.py:286-313 with their weights, under the names Keras gives them in a fresh session:
        the stride-2 layers are preceded by a ZeroPadding2D when ``fml_compatible`` (net.py:229-232)."""
        names = ["input_1"]
        fml = self._cfg.is_fml_compatible()
        if fml:
            names.append("zero_padding2d_1")
        names += ["separable_conv2d_1", "separable_conv2d_2"]
        if fml:
            names.append("zero_padding2d_2")
        names += ["separable_conv2d_3"] + [f"conv2d_{i}" for i in range(1, 8)]
        by_layer = {}
        for n, w in zip(_KERAS_NAMES, weights):
            by_layer.setdefault(n.split("/")[0], []).append((n + ":0", w))
        return [(n, by_layer.get(n, [])) for n in names]

    def save_weights(self, filepath, overwrite=True):
        """``model.save_weights`` (net.py:423): a Keras HDF5 weight file (``hdf5.write_keras_weights``; h5py is not
        needed) for ``.h5`` / ``.hdf5`` paths, else an ``.npz`` keyed by the Keras weight names in ``get_weights()`` order."""
        if not overwrite and os.path.exists(filepath):
            raise IOError(f"{filepath} exists")
        if str(filepath).endswith((".h5", ".hdf5", ".keras.h5")):
            from . import hdf5
            hdf5.write_keras_weights(filepath, self._keras_layers(self.get_weights()))
            return
        arrays = {f"{i:02d}:{n}": w for i, (n, w) in enumerate(zip(_KERAS_NAMES, self.get_weights()))}
        with open(filepath, "wb") as f:
            np.savez(f, **arrays)

    def load_weights(self, filepath, by_name=False):
        """``model.load_weights`` / the weight part of ``keras.models.load_model`` (net.py:424, 477): reads the
        HDF5 files Keras writes (weights-only or full ``model.save`` files) as well as this class's ``.npz``."""
        with open(filepath, "rb") as f:
            magic = f.read(8)
        if magic.startswith(b"\x89HDF"):
            from . import hdf5
            arrays, names = hdf5.read_keras_weights(filepath)
            if len(arrays) != len(_KERAS_NAMES):
                raise ValueError(f"{filepath} holds {len(arrays)} weight arrays, this architecture has {len(_KERAS_NAMES)} "
                                 "(net.py:286-313)")
            self.set_weights(arrays)
            return
        with np.load(filepath) as z:
            keys = sorted(z.files)
            self.set_weights([z[k] for k in keys])

    def save(self, filepath, overwrite=True, include_optimizer=True):
        """``model.save`` (net.py:419-420, 427): weights under ``model_weights`` of a full-model HDF5 file.  The
        ``model_config`` attribute names the architecture options only (Keras rebuilds the graph from its own JSON,
        which this class does not emit), optimizer state is not stored: ``NetManager.load_model`` rebuilds the model
        from ``config.pkl`` and loads the weights, which is all the reference's inference path needs."""
        if not overwrite and os.path.exists(filepath):
            raise IOError(f"{filepath} exists")
        if str(filepath).endswith((".h5", ".hdf5")):
            import json
            from . import hdf5
            cfg = json.dumps({"class_name": "Model", "config": {"name": self.name, "builder": "ubdvss_b200.net.B200Model",
                                                                "grey": self._cfg.is_grey(), "fml_compatible": self._cfg.is_fml_compatible(),
                                                                "n_classes": self.n_classes}})
            hdf5.write_keras_weights(filepath, self._keras_layers(self.get_weights()), full_model=True, model_config=cfg)
            return
        self.save_weights(filepath, overwrite)

    def summary(self, line_length=None, positions=None, print_fn=print):
        print_fn(f'Model: "{self.name}" (libubd.so, precision={self._engine.precision})')
        for n, s in zip(_KERAS_NAMES, weight_shapes(self._cfg.is_grey(), self.n_classes)):
            print_fn(f"  {n:44s} {str(s):18s} {int(np.prod(s)):8d}")
        print_fn(f"Total params: {self.count_params():,}")

    # ---- training (train.py:110-112, 176-188)
    def compile(self, optimizer, loss=None, metrics=None, **kwargs):
        from . import losses
        if not hasattr(optimizer, "lr"):
            raise ValueError("optimizer must expose .lr/.beta_1/.beta_2/.epsilon (ubdvss_b200.net.Adam)")
        self._optimizer = optimizer
        if loss is None or loss is losses.detection_loss or getattr(loss, "__name__", "") == "detection_loss":
            self._classification_loss = False
        elif loss is losses.detection_and_classification_loss or getattr(loss, "__name__", "") == "detection_and_classification_loss":
            if not self.n_classes:
                raise ValueError("classification loss needs a model with a class head")
            self._classification_loss = True
        else:
            raise ValueError("loss must come from ubdvss_b200.losses.get_loss (losses.py:20-24)")
        if self.n_classes and not self._classification_loss:
            raise ValueError("a model with a class head trains with detection_and_classification_loss (train.py:111)")
        # metrics: tokens from keras_metrics.get_all_metrics / losses.get_losses (train.py:112); without them the
        # loss components are returned under their own names
        self._metric_fns = list(metrics) if metrics else None
        if self._metric_fns is not None:
            for m in self._metric_fns:
                if not (callable(m) and hasattr(m, "__name__") and hasattr(m, "_fn")):
                    raise ValueError("metrics must come from ubdvss_b200.keras_metrics.get_all_metrics / losses.get_losses")
            self.metrics_names = ["loss"] + [m.__name__ for m in self._metric_fns]
        else:
            self.metrics_names = ["loss", "positive_loss", "negative_loss", "hard_negative_loss"] + \
                                 (["classification_loss"] if self._classification_loss else [])

    def set_distributed(self, enabled=True, native=True):
        """Data-parallel training: every rank keeps identical weights and the flat gradient buffer is summed over
        the ranks between backward and Adam.  ``native`` (default): the library's own NCCL communicator
        (``ubd_comm_init`` / ``ubd_allreduce_grads``); ``torch.distributed`` only carries the 128-byte unique id.
        ``native=False``: all-reduce through ``torch.distributed`` on the raw device buffer."""
        if not enabled:
            if getattr(self, "_native_comm", False):
                self._engine.comm_destroy()          # train_update would otherwise keep exchanging gradients
            self._dist = None
            self._native_comm = False
            return
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self._dist = dist
        self._native_comm = bool(native)
        if native:
            from .engine import Engine
            box = [Engine.comm_unique_id() if dist.get_rank() == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self._engine.comm_init(box[0], dist.get_rank(), dist.get_world_size())

    def _allreduce_grads(self):
        if getattr(self, "_native_comm", False):
            self._engine.allreduce_grads()
            return 1.0 / self._dist.get_world_size()
        import torch
        ptr, n = self._engine.grad_buffer()

        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        from .parallel import allreduce_mean_
        t = torch.as_tensor(d, device=f"cuda:{self.device}")
        scale = allreduce_mean_(t, self._dist)
        torch.cuda.current_stream(self.device).synchronize()
        return scale

    def train_on_batch(self, x, y, sample_weight=None, class_weight=None, preprocessing=None):
        """One optimizer step; returns ``[loss, positive, negative, hard_negative(, classification)]``."""
        if self._optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        o = self._optimizer
        if self._dist is None or getattr(self, "_native_comm", False):
            # step, gradient exchange inside the library and Adam queued back to back: one host synchronisation
            parts = self._engine.train_update(x, y, _PREPROC_CODE[preprocessing], o.lr, o.beta_1, o.beta_2, o.epsilon)
            return self._batch_outputs(parts)
        parts = self._engine.train_step(x, y, _PREPROC_CODE[preprocessing])
        scale = self._allreduce_grads()
        self._engine.adam_step(o.lr, o.beta_1, o.beta_2, o.epsilon, scale)
        return self._batch_outputs(parts)

    def _batch_outputs(self, parts):
        """``[loss, *metrics]`` of the batch the engine has just evaluated."""
        if getattr(self, "_metric_fns", None) is not None:
            counts = self._engine.metric_counts()
            return [float(parts[0])] + [m(counts, parts) for m in self._metric_fns]
        out = [float(parts[0]), float(parts[1]), float(parts[2]), float(parts[3])]
        if self._classification_loss:
            out.append(float(parts[4]))
        return out

    def test_on_batch(self, x, y, sample_weight=None, preprocessing=None):
        logits = self.predict(x, preprocessing=preprocessing)
        parts, _ = self._engine.loss(logits, y)
        return self._batch_outputs(parts)

    def fit_generator(self, generator, steps_per_epoch=None, epochs=1, verbose=1, callbacks=None,
                      validation_data=None, validation_steps=None, max_queue_size=10, workers=0,
                      use_multiprocessing=False, shuffle=True, initial_epoch=0, preprocessing=None, **kwargs):
        """The loop ``train.py:176-188`` drives: ``steps_per_epoch`` x ``train_on_batch`` per epoch, then
        validation and ``on_epoch_end(epoch, logs)`` of every callback (callbacks read ``self.model``,
        keras_callbacks.py:77)."""
        if steps_per_epoch is None:
            raise ValueError("steps_per_epoch is required for a generator")
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            else:
                cb.model = self
        history = {}
        for cb in callbacks:
            getattr(cb, "on_train_begin", lambda logs=None: None)({})
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                getattr(cb, "on_epoch_begin", lambda e, logs=None: None)(epoch, {})
            sums = np.zeros(len(self.metrics_names))
            for step in range(steps_per_epoch):
                batch = next(generator)
                vals = self.train_on_batch(batch[0], batch[1], preprocessing=preprocessing)
                sums += np.asarray(vals)
                for cb in callbacks:
                    getattr(cb, "on_batch_end", lambda b, logs=None: None)(step, dict(zip(self.metrics_names, vals)))
            logs = dict(zip(self.metrics_names, (sums / steps_per_epoch).tolist()))
            if validation_data is not None and validation_steps:
                vs = np.zeros(len(self.metrics_names))
                for _ in range(validation_steps):
                    vb = next(validation_data)
                    vs += np.asarray(self.test_on_batch(vb[0], vb[1], preprocessing=preprocessing))
                logs.update({"val_" + k: v for k, v in zip(self.metrics_names, (vs / validation_steps).tolist())})
            for k, v in logs.items():
                history.setdefault(k, []).append(v)
            if verbose:
                logging.info("epoch %d: %s", epoch + 1, ", ".join(f"{k}={v:.5f}" for k, v in logs.items()))
            for cb in callbacks:
                getattr(cb, "on_epoch_end", lambda e, logs=None: None)(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            getattr(cb, "on_train_end", lambda logs=None: None)({})
        self.history = history
        return self


class NetManager:
    """Builds / saves / loads the model (net.py:255-494); file names as in net.py:260-263."""

    CURRENT_MODEL_FILENAME = "model.h5"
    INFERENCE_MODEL_FILENAME = "inference_model.h5"
    MODEL_WEIGHTS_FILENAME = "model_weights.h5"
    PICKLED_CONFIG_FILENAME = "config.pkl"

    def __init__(self, log_dir, net_config=None, device=0, precision="fp32"):
        self._log_dir = log_dir
        self._device, self._precision = device, precision
        if net_config is not None:
            self._net_config = net_config
        else:
            self.load_config()
        self._model = None

    def build_model(self):
        """net.py:273-314: the dilated-conv model; sets ``scale = 4`` in the config."""
        self._model = B200Model(self._net_config, device=self._device, precision=self._precision)
        self._net_config._scale = 4
        return self._net_config

    def get_keras_model(self):
        return self._model

    def save_model(self, step=None):
        """net.py:418-420: ``model{step:03d}.h5`` and ``model.h5`` (plus the config next to them)."""
        os.makedirs(self._log_dir, exist_ok=True)
        if step is not None:
            self._model.save(os.path.join(self._log_dir, "model{:03d}.h5".format(step)))
        self._model.save(os.path.join(self._log_dir, self.CURRENT_MODEL_FILENAME))
        self.save_config()

    def save_inference(self):
        """net.py:422-427: weights-only file, then the rebuilt model without optimizer state."""
        os.makedirs(self._log_dir, exist_ok=True)
        weights_path = os.path.join(self._log_dir, self.MODEL_WEIGHTS_FILENAME)
        self._model.save_weights(weights_path)
        self.build_model()
        self._model.load_weights(weights_path)
        self._model.save(os.path.join(self._log_dir, self.INFERENCE_MODEL_FILENAME), include_optimizer=False)
        self.save_config()

    def load_another_model(self, another_log_dir):
        """net.py:429-441: take the model of another log dir (same architecture options); the options that do not
        change the graph stay those of this manager.  Returns the merged config."""
        other = NetManager(another_log_dir, self._net_config, device=self._device, precision=self._precision)
        other.load_config()
        other.load_model()
        self._model = other._model
        self._net_config = NetConfig.from_others(other._net_config, self._net_config.get_side_multiple(),
                                                 self._net_config.get_max_side(), self._net_config.get_min_pixels_for_detection())
        return self._net_config

    def load_model(self, path_to_model=None):
        """net.py:443-466: ``inference_model.h5`` else ``model.h5`` of the log dir (or an explicit path, also relative
        to the log dir); ``FileNotFoundError`` when neither exists.  Reads Keras HDF5 files (``hdf5.py``)."""
        if path_to_model is not None:
            if not os.path.exists(path_to_model):
                path_to_model = os.path.join(self._log_dir, path_to_model)
            return self._load_model(path_to_model)
        candidates = [self.INFERENCE_MODEL_FILENAME, self.CURRENT_MODEL_FILENAME]
        for fname in candidates:
            p = os.path.join(self._log_dir, fname)
            if os.path.exists(p):
                return self._load_model(p)
        raise FileNotFoundError(f"Model not found in dir {self._log_dir}. Must contain at least one of the following files {candidates}")

    def _load_model(self, path_to_model):
        assert os.path.exists(path_to_model), f"model fname does not exist, {path_to_model}"
        logging.info(f"loading model from {path_to_model}")
        self.build_model()
        self._model.load_weights(path_to_model)
        return self._net_config

    def save_config(self):
        os.makedirs(self._log_dir, exist_ok=True)
        with open(os.path.join(self._log_dir, self.PICKLED_CONFIG_FILENAME), "wb") as f:
            pickle.dump(self._net_config, f)

    def load_config(self):
        with open(os.path.join(self._log_dir, self.PICKLED_CONFIG_FILENAME), "rb") as f:
            self._net_config = _ConfigUnpickler(io.BytesIO(f.read())).load()
        return self._net_config
