"""Drop-in for `src.sdav.network.SDAV.SDAV` (reference src/sdav/network/SDAV.py).

Same attributes and method names as the reference class; `transform(x)` - the encoder forward, the hot path - runs on
the B200 through libdlc (patches -> 5 fused GEMM+bias+sigmoid tcgen05 kernels) and returns the same flat float64
[B*30, 2500] array (SDAV.py:163, 293-302). Weights are explicit and persistent here (the reference re-initialises
or re-restores them inside every call, SDAV.py:232-240): seeded N(0,1) like `tf.random_normal` by default, or loaded
from / saved to an .npz (`w{l}_e [in,out]`, `b{l}_e [out]`, `b{l}_d [in]`, float64) or the reference's own TensorFlow
checkpoints (`tf_checkpoint.py`: variables `Variable` ... `Variable_14`, SDAV.py:188-217, 228-240).
Training (`fit`, `fit_dataset`: the reference's greedy layer schedule, SDAV.py:242-288) runs every SGD step on the
B200 through `training.DaeStackTrainer` (tcgen05 GEMMs + the dlc_train_* kernels).
"""
import glob
import logging
import os

import numpy as np

from . import input_parser, tf_checkpoint


class SDAV:
    def __init__(self, verbosity=logging.WARNING, weights_path=None, seed=0, precision="fp16x2", train_path=None,
                 input_shape=None, hidden_units=None):
        self._configure_logging(verbosity)
        self._set_train_path(train_path)
        self._define_params()
        if input_shape is not None:          # extension: other geometries than the reference's fixed 30 x 1681 -> 5 x 2500
            self.input_shape = list(input_shape)
        if hidden_units is not None:
            self.hidden_units = list(hidden_units)
        self.losses = []
        self.precision = precision
        self._encoder = None
        self._weights = None
        self._biases = None
        self._dec_biases = None
        self.global_step = 0
        if weights_path is None and os.path.isdir(self.checkpoints_path):
            # SDAV._load_or_init_session (SDAV.py:232-240): restore the latest checkpoint if the directory has one
            cand = os.path.join(self.checkpoints_path, "sdav_weights.npz")
            if os.path.exists(cand):
                weights_path = cand
            elif tf_checkpoint.latest_checkpoint(self.checkpoints_path):
                weights_path = self.checkpoints_path
        if weights_path is not None:
            self.load_weights(weights_path)
        else:
            self.init_weights(seed)
        logging.info('Done initializing sdav')

    # ---- parameters (SDAV.py:30-39)
    def _define_params(self):
        self.input_shape = [30, 1681]
        self.hidden_units = [2500, 2500, 2500, 2500, 2500]
        self.default_batch_size = 10
        self.sparse_level = 0.05
        self.sparse_penalty = 1.0
        self.consecutive_penalty = 0.2
        self.learning_rate = 0.1
        self.epochs = 100
        self.corruption_level = 0.3

    def _set_train_path(self, train_path):
        # the reference derives this from the location of a directory named 'deepLoopCloser' (PathUtils.py:1-11);
        # here it is an explicit argument with a local default and nothing is created on disk until save_weights().
        self.train_path = train_path or os.path.join(os.getcwd(), "training", "sdav")
        self.checkpoints_path = self.train_path + '/checkpoints'
        self.log_path = self.train_path + '/log'

    def _configure_logging(self, verbosity):
        self.logger = logging.getLogger()
        self.logger.setLevel(verbosity)

    @property
    def dims(self):
        return [self.input_shape[1]] + list(self.hidden_units)

    # ---- weights
    def init_weights(self, seed=0):
        """tf.random_normal weights (sigma = 1), zero biases (SDAV.py:189-217), seeded."""
        rng = np.random.default_rng(seed)
        d = self.dims
        self.set_weights([rng.standard_normal((k, n)) for k, n in zip(d[:-1], d[1:])], [np.zeros(n) for n in d[1:]])

    def set_weights(self, weights, biases, decoder_biases=None):
        d = self.dims
        if len(weights) != len(d) - 1 or len(biases) != len(d) - 1:
            raise ValueError("expected %d weight matrices and biases" % (len(d) - 1))
        self._weights = [np.ascontiguousarray(w, dtype=np.float64) for w in weights]
        self._biases = [np.ascontiguousarray(b, dtype=np.float64) for b in biases]
        if decoder_biases is None:
            decoder_biases = [np.zeros(k) for k in d[:-1]]           # tf.zeros (SDAV.py:193, 199, ...)
        self._dec_biases = [np.ascontiguousarray(b, dtype=np.float64) for b in decoder_biases]
        for l, (w, b, bd) in enumerate(zip(self._weights, self._biases, self._dec_biases)):
            if w.shape != (d[l], d[l + 1]) or b.shape != (d[l + 1],) or bd.shape != (d[l],):
                raise ValueError("layer %d: expected W %s, b %s, decoder b %s" % (l, (d[l], d[l + 1]), (d[l + 1],), (d[l],)))
        self._encoder = None  # re-packed lazily on the device

    def load_weights(self, path):
        """`path`: an .npz written by save_weights, a TensorFlow checkpoint prefix, or a directory holding the
        reference's `checkpoint` state file (the latest checkpoint is restored, like SDAV.py:232-236)."""
        n = len(self.hidden_units)
        if path.endswith(".npz"):
            z = np.load(path)
            dec = [z["b%d_d" % l] for l in range(n)] if "b0_d" in z else None
            self.set_weights([z["w%d_e" % l] for l in range(n)], [z["b%d_e" % l] for l in range(n)], dec)
            self.global_step = int(z["global_step"]) if "global_step" in z else 0
            return
        prefix = path
        if os.path.isdir(path):
            prefix = tf_checkpoint.latest_checkpoint(path)
            if prefix is None:
                raise FileNotFoundError("no TensorFlow checkpoint state file in %s" % path)
        logging.info('Restoring session from %s' % prefix)
        tensors = tf_checkpoint.load_checkpoint(prefix)
        names = tf_checkpoint.sdav_variable_names(n)
        missing = [nm for trio in names for nm in trio if nm not in tensors]
        if missing:
            raise KeyError("checkpoint %s lacks the SDAV variables %s (has %s)" % (prefix, missing, sorted(tensors)))
        self.set_weights([tensors[w] for w, _, _ in names], [tensors[b] for _, b, _ in names],
                         [tensors[bd] for _, _, bd in names])
        self.global_step = int(tensors["global_step"]) if "global_step" in tensors else 0

    def save_weights(self, path=None, fmt=None):
        """fmt "npz" (default for *.npz paths) or "tf": a TensorFlow V2 checkpoint the reference can restore
        (`<checkpoints>/checkpoint_file-<global_step>` + the `checkpoint` state file, SDAV.py:228-230, 273-275)."""
        if fmt is None:
            fmt = "tf" if (path is not None and not path.endswith(".npz")) else "npz"
        step = int(getattr(self, "global_step", 0))
        if fmt == "tf":
            prefix = path or "%s-%d" % (os.path.join(self.checkpoints_path, "checkpoint_file"), step)
            tensors = {"global_step": np.array(step, dtype=np.int32)}
            for (wn, bn, dn), w, b, bd in zip(tf_checkpoint.sdav_variable_names(len(self.hidden_units)),
                                              self._weights, self._biases, self._dec_biases):
                tensors[wn], tensors[bn], tensors[dn] = w, b, bd
            return tf_checkpoint.save_checkpoint(prefix, tensors)
        path = path or os.path.join(self.checkpoints_path, "sdav_weights.npz")
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        arrays = {"global_step": np.array(step, dtype=np.int32)}
        for l, (w, b, bd) in enumerate(zip(self._weights, self._biases, self._dec_biases)):
            arrays["w%d_e" % l] = w
            arrays["b%d_e" % l] = b
            arrays["b%d_d" % l] = bd
        np.savez(path, **arrays)
        return path

    def _get_encoder(self):
        if self._encoder is None:
            from . import _cuda, ops
            _cuda.require_cuda()
            enc = ops.SdaEncoder(self.dims, self.precision)
            for l, (w, b) in enumerate(zip(self._weights, self._biases)):
                enc.set_layer(l, w, b)
            self._encoder = enc
        return self._encoder

    # ---- shapes (SDAV.py:165-169, 290-291)
    def get_layer_input_shape(self, layer_n):
        if layer_n == 0:
            return self.input_shape
        return [self.input_shape[0], self.hidden_units[layer_n - 1]]

    def get_layers_input_shapes(self):
        return list(map(self.get_layer_input_shape, range(1, 6)))

    # ---- the hot path
    def transform(self, x):
        """x float64 [B, 30, 1681] -> float64 [B*30, 2500] (flat, like the reference's `_h4`)."""
        import torch
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 3 or list(x.shape[1:]) != self.input_shape:
            raise ValueError("expected input of shape [B, %d, %d], got %s" % (self.input_shape[0], self.input_shape[1], x.shape))
        enc = self._get_encoder()
        flat = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, x.shape[-1]))).cuda()
        out = enc.encode(flat)
        return out.to(torch.float64).cpu().numpy()

    def transform_dataset(self, file_pattern, key_points=None):
        """Encode every image matched by `file_pattern` (sorted) -> float64 [N, 30, 2500]. The reference's version
        (SDAV.py:309-318) cannot run as written; this is its evident intent, and what
        create_similarity_matrix.py:27 calls as `transform_all`. `key_points`: optional [N, 30, 2] (x, y) array or a
        callable image -> [30, 2]; without it the SURF detector of CvInputParser is required."""
        files = sorted(glob.glob(file_pattern))
        if len(files) == 0:
            logging.getLogger().error("Specified dataset is empty or could not find dataset")
            return np.zeros((0, self.input_shape[0], self.hidden_units[-1]))
        patch_size = int(round(np.sqrt(self.input_shape[1])))
        parser = input_parser.CvInputParser(self.input_shape[0], patch_size)
        frames = []
        for i, f in enumerate(files):
            kp = key_points(i, f) if callable(key_points) else (None if key_points is None else key_points[i])
            frames.append(parser.parse_from_path(f, key_points=kp))
        x = np.stack(frames)
        return self.transform(x).reshape(len(files), self.input_shape[0], self.hidden_units[-1])

    transform_all = transform_dataset  # name used by src/sdav/create_similarity_matrix.py:27

    def get_dataset(self, file_pattern):
        """Reference: a tf.data generator dataset of parsed frames (SDAV.py:219-221). Here: the sorted file list."""
        return sorted(glob.glob(file_pattern))

    # ---- training: the reference's greedy schedule, each step on the B200
    def _get_trainer(self):
        from .training import DaeStackTrainer
        tr = DaeStackTrainer(self.dims, patches=self.input_shape[0], sparse_level=self.sparse_level,
                             sparse_penalty=self.sparse_penalty, consecutive_penalty=self.consecutive_penalty,
                             learning_rate=self.learning_rate)
        tr.set_weights(self._weights, self._biases, self._dec_biases)
        tr.global_step = int(self.global_step)
        return tr

    def _adopt(self, trainer):
        ws, bs, bds = trainer.get_weights()
        self.set_weights(ws, bs, bds)
        self.global_step = int(trainer.global_step)

    def _fit_batch(self, trainer, layer, batch, generator=None, log_batch=None):
        import torch
        xd = torch.from_numpy(np.ascontiguousarray(batch, dtype=np.float32)).cuda()
        loss = None
        for step in range(self.epochs):
            masks = trainer.sdav_masks(layer, self.corruption_level, generator)      # redrawn on every run (:34-38)
            if self.epochs >= 4:   # the step is launch-bound at the reference's batch size: replay it as a CUDA graph
                graphed = trainer.cached_graphed_step(xd, layer, masks)             # captured once per (shape, layer)
                loss = graphed(xd, masks)
            else:
                loss = trainer.step(xd, layer, masks)
            if log_batch is not None and self.logger.isEnabledFor(logging.INFO):
                logging.info('    Layer:%d Batch:%d fit, Epoch:%d/%d, Loss:%s' % (layer, log_batch, step + 1,
                                                                                self.epochs, float(loss.item())))
        return None if loss is None else float(loss.item())

    def fit(self, x, seed=None):
        """SDAV.fit (SDAV.py:277-288): for each of the five layers, `epochs` SGD steps of train_steps[i] on the batch
        x [B, 30, 1681]. Returns the last loss of every layer."""
        import torch
        x = np.asarray(x, dtype=np.float64)
        if x.ndim != 3 or list(x.shape[1:]) != self.input_shape:
            raise ValueError("expected input of shape [B, %d, %d], got %s" % (self.input_shape[0], self.input_shape[1], x.shape))
        gen = None
        if seed is not None:
            gen = torch.Generator(device="cuda")
            gen.manual_seed(seed)
        trainer = self._get_trainer()
        losses = [self._fit_batch(trainer, i, x, gen) for i in range(len(self.hidden_units))]
        self._adopt(trainer)
        return losses

    def fit_dataset(self, dataset, key_points=None, seed=None, save=True):
        """SDAV.fit_dataset (SDAV.py:242-275): for each layer, walk the dataset in batches of `default_batch_size`
        frames and run `epochs` steps of train_steps[i] on every batch; a checkpoint is written after each layer
        (TensorFlow format, `<checkpoints>/checkpoint_file-<global_step>`). `dataset`: the file list returned by
        get_dataset(), or an iterable of parsed frames [30, 1681]; `key_points` as in transform_dataset."""
        import torch
        frames = self._materialise(dataset, key_points)
        gen = None
        if seed is not None:
            gen = torch.Generator(device="cuda")
            gen.manual_seed(seed)
        trainer = self._get_trainer()
        bs = self.default_batch_size
        for i in range(len(self.hidden_units)):
            logging.info('Fitting layer %d' % i)
            for batch_n, s in enumerate(range(0, len(frames), bs)):
                batch = frames[s:s + bs]
                if len(batch) < 2:       # the consecutive-frame term of a single frame is a mean over nothing (NaN)
                    logging.warning("Ignored a trailing batch of one frame")
                    continue
                self._fit_batch(trainer, i, batch, gen, log_batch=batch_n)
            self._adopt(trainer)
            if save:
                logging.info('Saving trained params to %s with global_step %s' % (self.checkpoints_path, self.global_step))
                self.save_weights(fmt="tf")

    def _materialise(self, dataset, key_points):
        if len(dataset) and isinstance(dataset[0], str):
            patch_size = int(round(np.sqrt(self.input_shape[1])))
            parser = input_parser.CvInputParser(self.input_shape[0], patch_size)
            out = []
            for i, f in enumerate(dataset):
                kp = key_points(i, f) if callable(key_points) else (None if key_points is None else key_points[i])
                out.append(parser.parse_from_path(f, key_points=kp))
            return np.stack(out)
        return np.asarray(dataset, dtype=np.float64)
