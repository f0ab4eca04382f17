"""Drop-ins for `src.sdav.network.DenoisingAutoencoderVariant.DA` and
`src.sdav.network.StackedDenoisingAutoencoderVariants.SDA` (reference files of the same names).
`DA.transform(x[30, in]) -> [30, hidden]` = sigmoid(x w0 + b0) (DenoisingAutoencoderVariant.py:116-119, 254-259) runs
as one fused tcgen05 kernel; `SDA` chains the layers (StackedDenoisingAutoencoderVariants.py:73-83). Constructor
arguments are the reference's (the `graph` argument is accepted and ignored). `fit` / `fit_dataset` run the reference's
training schedule with every SGD step on the B200 (`training.DaeStackTrainer`)."""
import logging

import numpy as np


def _check_common(sparse_level, sparse_penalty, consecutive_penalty, batch_size, learning_rate, epochs,
                  corruption_level):
    # the reference validates with py_v8n (DenoisingAutoencoderVariant.py:67-90); same constraints, plain Python
    for name, v in (("learning_rate", learning_rate), ("sparse_level", sparse_level)):
        if not isinstance(v, float) or v <= 0:
            raise ValueError("%s must be a positive float" % name)
    for name, v in (("sparse_penalty", sparse_penalty), ("consecutive_penalty", consecutive_penalty),
                    ("corruption_level", corruption_level)):
        if not isinstance(v, (float, int)) or not (0 <= v <= 1):
            raise ValueError("%s must be in [0, 1]" % name)
    for name, v in (("batch_size", batch_size), ("epochs", epochs)):
        if not isinstance(v, int) or v <= 0:
            raise ValueError("%s must be a positive int" % name)


class DA:
    def __init__(self, input_shape, hidden_units, sparse_level=0.05, sparse_penalty=1.0, consecutive_penalty=0.2,
                 batch_size=10, learning_rate=0.1, epochs=100, layer_n=0, corruption_level=0.3, graph=None,
                 seed=0, precision="fp16x2"):
        _check_common(sparse_level, sparse_penalty, consecutive_penalty, batch_size, learning_rate, epochs,
                      corruption_level)
        if (not isinstance(input_shape, (list, tuple)) or len(input_shape) != 2 or
                not all(isinstance(v, int) and v > 0 for v in input_shape)):
            raise ValueError("input_shape must be a list of two positive ints")
        if not isinstance(hidden_units, int) or hidden_units <= 0:
            raise ValueError("hidden_units must be a positive int")
        self.input_shape = list(input_shape)
        self.hidden_units = hidden_units
        self.sparse_level = sparse_level
        self.sparse_penalty = sparse_penalty
        self.consecutive_penalty = consecutive_penalty
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.epochs = epochs
        self.corruption_level = corruption_level
        self.layer_n = layer_n
        self.precision = precision
        self._encoder = None
        rng = np.random.default_rng(seed + layer_n)
        # encoder_weights ~ N(0,1), encoder_biases = 0 (DenoisingAutoencoderVariant.py:94-97)
        self.set_weights(rng.standard_normal((self.input_shape[1], hidden_units)), np.zeros(hidden_units))

    def set_weights(self, w0, b0, b1=None):
        w0 = np.ascontiguousarray(w0, dtype=np.float64)
        b0 = np.ascontiguousarray(b0, dtype=np.float64)
        b1 = np.zeros(self.input_shape[1]) if b1 is None else np.ascontiguousarray(b1, dtype=np.float64)
        if (w0.shape != (self.input_shape[1], self.hidden_units) or b0.shape != (self.hidden_units,) or
                b1.shape != (self.input_shape[1],)):
            raise ValueError("expected w0 %s, b0 %s and b1 %s" % ((self.input_shape[1], self.hidden_units),
                                                                 (self.hidden_units,), (self.input_shape[1],)))
        self._w0, self._b0, self._b1 = w0, b0, b1
        self._encoder = None

    # ---- the reference's checkpoints (DenoisingAutoencoderVariant.py:160-174): <train>/checkpoints/layer_<n>_/
    def load_checkpoint(self, path):
        """Restore encoder_variables/encoder_weights, encoder_biases and decoder_variables/decoder_biases from a
        TensorFlow checkpoint prefix or a directory holding the `checkpoint` state file."""
        import os

        from . import tf_checkpoint
        prefix = tf_checkpoint.latest_checkpoint(path) if os.path.isdir(path) else path
        if prefix is None:
            raise FileNotFoundError("no TensorFlow checkpoint state file in %s" % path)
        t = tf_checkpoint.load_checkpoint(prefix)
        wn, bn, dn = tf_checkpoint.DA_VARIABLE_NAMES
        self.set_weights(t[wn], t[bn], t[dn])
        self.global_step = int(t["global_step"]) if "global_step" in t else 0

    def save_checkpoint(self, prefix):
        import numpy as _np

        from . import tf_checkpoint
        wn, bn, dn = tf_checkpoint.DA_VARIABLE_NAMES
        return tf_checkpoint.save_checkpoint(prefix, {wn: self._w0, bn: self._b0, dn: self._b1,
                                                      "global_step": _np.array(int(getattr(self, "global_step", 0)),
                                                                               dtype=_np.int32)})

    def transform(self, x, batch_n=-1):
        import torch

        from . import _cuda, ops
        logging.info("  Layer %d: transform" % self.layer_n)
        x = np.asarray(x, dtype=np.float64)
        if list(x.shape) != self.input_shape:
            raise ValueError("expected input of shape %s, got %s" % (self.input_shape, list(x.shape)))
        if self._encoder is None:
            _cuda.require_cuda()
            self._encoder = ops.SdaEncoder([self.input_shape[1], self.hidden_units], self.precision)
            self._encoder.set_layer(0, self._w0, self._b0)
        out = self._encoder.encode(torch.from_numpy(np.ascontiguousarray(x)).cuda())
        return out.to(torch.float64).cpu().numpy()

    # ---- training (DenoisingAutoencoderVariant.py:103-158, 182-243): every SGD step on the B200
    def fit(self, file_pattern, key_points=None):
        """DA.fit (:204-208): parse the images matched by `file_pattern` into frames and fit on them."""
        import glob

        from . import input_parser
        logging.info("  Layer:%d fit" % self.layer_n)
        patch_size = int(round(np.sqrt(self.input_shape[1])))
        parser = input_parser.CvInputParser(self.input_shape[0], patch_size)
        frames = []
        for i, f in enumerate(sorted(glob.glob(file_pattern))):
            kp = key_points(i, f) if callable(key_points) else (None if key_points is None else key_points[i])
            frames.append(parser.parse_from_path(f, key_points=kp))
        return self.fit_dataset(frames)

    def fit_dataset(self, dataset, seed=None):
        """DA.fit_dataset (:210-243): batches of `batch_size` frames [30, in]; `epochs` steps of the train_step on each;
        a trailing batch smaller than batch_size is ignored with the reference's warning. The corruption masks are
        drawn once per instance, like the graph constants of _corrupt_tensor (:182-202). Returns the last loss."""
        import torch

        from .training import DaeStackTrainer
        frames = [np.asarray(f, dtype=np.float64) for f in dataset]
        trainer = DaeStackTrainer([self.input_shape[1], self.hidden_units], patches=self.input_shape[0],
                                  sparse_level=self.sparse_level, sparse_penalty=self.sparse_penalty,
                                  consecutive_penalty=self.consecutive_penalty, learning_rate=self.learning_rate)
        trainer.set_weights([self._w0], [self._b0], [self._b1])
        trainer.global_step = int(getattr(self, "global_step", 0))
        rows = self.batch_size * self.input_shape[0]
        # the masks are graph constants in the reference (drawn once per instance); here once per (seed, rows): a
        # later call with another seed or batch size draws its own instead of silently reusing the first ones
        mask_key = (seed, rows)
        if getattr(self, "_masks", None) is None or getattr(self, "_masks_key", None) != mask_key:
            gen = None
            if seed is not None:
                gen = torch.Generator(device="cuda")
                gen.manual_seed(seed)
            self._masks = trainer.da_masks(rows, self.corruption_level, gen)
            self._masks_key = mask_key
        zm, om = self._masks
        loss = None
        for batch_n, s in enumerate(range(0, len(frames), self.batch_size)):
            batch = frames[s:s + self.batch_size]
            if len(batch) != self.batch_size:
                logging.warning("Ignored last batch because it was smaller than the specified batch size. To avoid this "
                                "choose a batch size that is a factor of the dataset size.")
                break
            xd = torch.from_numpy(np.ascontiguousarray(np.stack(batch), dtype=np.float32)).cuda()
            for step in range(self.epochs):
                if self.epochs >= 4:   # launch-bound step: captured ONCE per step shape, replayed as a CUDA graph
                    graphed = trainer.cached_graphed_step(xd, 0, [zm], [om], mask_rows=rows, da_mode=True)
                    loss = graphed(xd)
                else:
                    loss = trainer.step(xd, 0, [zm], [om], mask_rows=rows, da_mode=True)
                if logging.getLogger().isEnabledFor(logging.INFO):
                    logging.info('    Layer:%d Batch:%d fit, Epoch:%d/%d, Loss:%s' % (self.layer_n, batch_n, step + 1,
                                                                                    self.epochs, float(loss.item())))
        ws, bs, bds = trainer.get_weights()
        self.set_weights(ws[0], bs[0], bds[0])
        self.global_step = int(trainer.global_step)
        return None if loss is None else float(loss.item())


class SDA:
    def __init__(self, input_shape, hidden_units, sparse_level=0.05, sparse_penalty=1, consecutive_penalty=0.2,
                 batch_size=10, learning_rate=0.1, epochs=100, corruption_level=0.3, graph=None, seed=0,
                 precision="fp16x2"):
        if not isinstance(hidden_units, (list, tuple)) or len(hidden_units) < 2:
            raise ValueError("hidden_units must list at least two layer widths")  # min_length(2), SDA :56
        self.input_shape = list(input_shape)
        self.hidden_units = list(hidden_units)
        self.sparse_level = sparse_level
        self.sparse_penalty = sparse_penalty
        self.consecutive_penalty = consecutive_penalty
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.epochs = epochs
        self.corruption_level = corruption_level
        self._layers = []
        for i, h in enumerate(self.hidden_units):  # StackedDenoisingAutoencoderVariants.py:73-83
            shape = self.input_shape if i == 0 else [self.input_shape[0], self.hidden_units[i - 1]]
            self._layers.append(DA(shape, h, sparse_level=float(sparse_level), sparse_penalty=float(sparse_penalty),
                                   consecutive_penalty=float(consecutive_penalty), batch_size=batch_size,
                                   learning_rate=float(learning_rate), epochs=epochs, layer_n=i,
                                   corruption_level=float(corruption_level), seed=seed, precision=precision))

    def transform(self, x):
        """Chain the layers' transforms for one frame x [30, in] -> [30, hidden[-1]] (what SDA.fit feeds layer i+1
        with, StackedDenoisingAutoencoderVariants.py:96-100)."""
        for layer in self._layers:
            x = layer.transform(x)
        return x

    def fit(self, file_pattern, key_points=None):
        """SDA.fit (StackedDenoisingAutoencoderVariants.py:90-100): fit layer 0 on the parsed frames, then each next
        layer on the frames mapped through the previous layers' transform."""
        import glob

        from . import input_parser
        logging.info("Fit SDAV")
        patch_size = int(round(np.sqrt(self.input_shape[1])))
        parser = input_parser.CvInputParser(self.input_shape[0], patch_size)
        frames = []
        for i, f in enumerate(sorted(glob.glob(file_pattern))):
            kp = key_points(i, f) if callable(key_points) else (None if key_points is None else key_points[i])
            frames.append(parser.parse_from_path(f, key_points=kp))
        return self.fit_dataset(frames)

    def fit_dataset(self, frames):
        losses = []
        for i, layer in enumerate(self._layers):
            losses.append(layer.fit_dataset(frames))
            if i + 1 < len(self._layers):
                frames = [layer.transform(f) for f in frames]
        return losses
