"""Drop-in for `src.cnn_vtl.network.cnn_vtl.CnnVtl` (reference src/cnn_vtl/network/cnn_vtl.py).

transform(x[N,H,W,3]) -> int8 [N, M]: AlexNet conv1..conv5 (ungrouped, no LRN, conv5 linear, cnn_vtl.py:33-93), all
five conv outputs flattened + concatenated (:96-106), per-image min/max scaling to 0..255 and int8 cast (:109-116),
random column sub-sampling fixed per instance (:119-128). On the B200 every convolution is ONE tcgen05 kernel behind
a `dlc_cnnvtl` handle: conv2..conv5 are implicit GEMMs (im2col-mode TMA reads the NHWC activation tap by tap), and
bias + ReLU, the per-image min/max and the gather of the kept columns are fused into the epilogue, so neither an
im2col matrix nor the 546,944-wide float descriptor is ever written. `_forward_chunk_explicit` is the
explicit-im2col formulation built from the library's building blocks, kept as a cross-check.

Differences from the reference, all explicit: weights are an argument (the AlexNet blob is a git-LFS pointer in the
reference tree) - a dict {name: (W[kh,kw,cin,cout], b[cout])}, a path to a `bvlc_alexnet.npy`-style file, or
`weights="synthetic"` for seeded He-scaled tensors; the column mask can be passed (`mask=` boolean [sum sizes] or
`keep_cols=`) or seeded (`seed=`); frames are processed in device chunks of `DEVICE_CHUNK` images to bound memory."""
import numpy as np

# (name, kh, kw, cin, cout, stride, padding, relu)
LAYERS = (
    ("conv1", 11, 11, 3, 96, 4, "valid", True),
    ("conv2", 5, 5, 96, 256, 1, "same", True),
    ("conv3", 3, 3, 256, 384, 1, "same", True),
    ("conv4", 3, 3, 384, 384, 1, "same", True),
    ("conv5", 3, 3, 384, 256, 1, "same", False),
)
POOL_AFTER = ("conv1", "conv2")


def compressed_size(value, compression):
    """src/utils/MathUtils.py:1-4."""
    return int(round(value * ((100 - compression) / 100)))


def _out_hw(h, w, k, stride, padding):
    if padding == "same":
        return -(-h // stride), -(-w // stride)
    return (h - k) // stride + 1, (w - k) // stride + 1


def _same_pad_before(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2  # TF puts the extra pixel after


def layer_geometry(height, width):
    """Per conv layer: (H_in, W_in, OH, OW, pad_t, pad_l) for an input of height x width."""
    geo = []
    h, w = height, width
    for name, kh, kw, cin, cout, stride, padding, relu in LAYERS:
        oh, ow = _out_hw(h, w, kh, stride, padding)
        pt = _same_pad_before(h, kh, stride) if padding == "same" else 0
        pl = _same_pad_before(w, kw, stride) if padding == "same" else 0
        geo.append((h, w, oh, ow, pt, pl))
        h, w = oh, ow
        if name in POOL_AFTER:
            h, w = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    return geo


def synthetic_weights(seed):
    rng = np.random.default_rng(seed)
    params = {}
    for name, kh, kw, cin, cout, *_ in LAYERS:
        params[name] = (rng.standard_normal((kh, kw, cin, cout)) * np.sqrt(2.0 / (kh * kw * cin)),
                        0.05 * rng.standard_normal(cout))
    return params


def _flat_fill(values, shape):
    """tf.constant_initializer semantics for a value list shorter than the variable: fill the remainder with the
    last value [TF1-doc] - this is what initialising UNGROUPED convs from grouped AlexNet tensors relies on
    (cnn_vtl.py:140-148)."""
    flat = np.asarray(values, dtype=np.float64).ravel()
    n = int(np.prod(shape))
    if flat.size > n:
        raise ValueError("too many elements (%d) for shape %s" % (flat.size, shape))
    if flat.size < n:
        flat = np.concatenate([flat, np.full(n - flat.size, flat[-1])])
    return flat.reshape(shape)


def load_alexnet_npy(path):
    raw = np.load(path, encoding="bytes", allow_pickle=True).item()
    params = {}
    for name, kh, kw, cin, cout, *_ in LAYERS:
        key = name if name in raw else name.encode()
        params[name] = (_flat_fill(raw[key][0], (kh, kw, cin, cout)), _flat_fill(raw[key][1], (cout,)))
    return params


class CnnVtl:
    def __init__(self, input_shape=(1, 224, 224, 3), batch_size=10, compress_factor=99.59, weights=None, mask=None,
                 keep_cols=None, seed=None, precision="fp16x2"):
        self.input_shape = input_shape
        self.batch_size = batch_size
        self.compress_factor = compress_factor
        self.precision = precision
        _, self._H, self._W, c = input_shape
        if c != 3:
            raise ValueError("input_shape must be [N, H, W, 3]")
        self._geo = layer_geometry(self._H, self._W)
        self.layer_sizes = [g[2] * g[3] * L[4] for g, L in zip(self._geo, LAYERS)]
        # ---- weights
        if weights is None:
            raise ValueError("CnnVtl needs weights=: a {name: (W, b)} dict, a path to bvlc_alexnet.npy, or "
                             "'synthetic' (the reference's pretrained blob is not shipped with its repository)")
        if isinstance(weights, str):
            weights = synthetic_weights(0 if seed is None else seed) if weights == "synthetic" else load_alexnet_npy(weights)
        self.params = {}
        for name, kh, kw, cin, cout, *_ in LAYERS:
            w, b = weights[name]
            w = np.ascontiguousarray(w, dtype=np.float64)
            b = np.ascontiguousarray(b, dtype=np.float64)
            if w.shape != (kh, kw, cin, cout) or b.shape != (cout,):
                raise ValueError("%s: expected W %s and b %s" % (name, (kh, kw, cin, cout), (cout,)))
            self.params[name] = (w, b)
        # ---- column mask (cnn_vtl.py:119-126: per layer, compressed_size indices drawn WITH replacement)
        total = int(np.sum(self.layer_sizes))
        if keep_cols is not None:
            self.keep_cols = np.unique(np.asarray(keep_cols, dtype=np.int64))
        elif mask is not None:
            mask = np.asarray(mask, dtype=bool)
            if mask.shape != (total,):
                raise ValueError("mask must have shape (%d,)" % total)
            self.keep_cols = np.flatnonzero(mask).astype(np.int64)
        else:
            rng = np.random.default_rng(seed) if seed is not None else np.random.default_rng()
            cols, start = [], 0
            for s in self.layer_sizes:
                cols.append(np.unique(rng.choice(np.arange(start, start + s), size=compressed_size(s, compress_factor))))
                start += s
            self.keep_cols = np.concatenate(cols).astype(np.int64)
        if self.keep_cols.size and (self.keep_cols.min() < 0 or self.keep_cols.max() >= total):
            raise ValueError("kept columns out of range")
        self._dev = None
        self._head = None

    # ---- fused path: one dlc_cnnvtl handle per instance
    def _fused_head(self):
        if self._head is None:
            from . import _cuda, ops
            _cuda.require_cuda()
            head = ops.CnnVtlHead(self._H, self._W, self.precision)
            for l, (name, *_r) in enumerate(LAYERS):
                head.set_conv(l, *self.params[name])
            if self.keep_cols.size:
                head.set_keep_cols(self.keep_cols)
            self._head = head
        return self._head

    def _forward_chunk(self, x):
        """x: CUDA tensor [n, H, W, 3] (uint8 / float32 / float64) -> int8 [n, M]."""
        import torch
        if x.dtype not in (torch.uint8, torch.float32, torch.float64):
            x = x.to(torch.float64)
        head = self._fused_head()
        if not self.keep_cols.size:
            return torch.empty((x.shape[0], 0), dtype=torch.int8, device=x.device)
        return head.forward(x.contiguous())

    def conv_outputs(self, x):
        """Diagnostics: float32 NHWC outputs of conv1..conv5 for a CUDA batch x (fused path)."""
        import torch
        head = self._fused_head()
        n = x.shape[0]
        outs = [torch.empty((n, g[2], g[3], L[4]), dtype=torch.float32, device=x.device)
                for g, L in zip(self._geo, LAYERS)]
        head.forward(x.contiguous(), layer_outputs=outs, quantise=False)
        return outs

    # ---- explicit-im2col path (cross-check): packed weight planes, biases, kept columns
    def _device_state(self):
        if self._dev is None:
            import torch

            from . import _cuda, ops
            _cuda.require_cuda()
            split = self.precision == "fp16x2"
            st = {"w": {}, "b": {}}
            for name, kh, kw, cin, cout, *_ in LAYERS:
                w, b = self.params[name]
                wt = torch.from_numpy(w.reshape(kh * kw * cin, cout)).cuda()
                st["w"][name] = ops.pack_weight_planes(wt, n_pad=cout, need_lo=split)
                st["b"][name] = torch.from_numpy(b.astype(np.float32)).cuda()
            st["keep"] = torch.from_numpy(self.keep_cols).cuda()
            self._dev = st
        return self._dev

    def _forward_chunk_explicit(self, x):
        """x: CUDA tensor [n, H, W, 3] (uint8 / float) -> int8 [n, M]."""
        import torch

        from . import ops
        st = self._device_state()
        split = self.precision == "fp16x2"
        n = x.shape[0]
        flat = x.reshape(n * self._H * self._W, 3).to(torch.float32).contiguous()
        hi, lo = ops.split_planes(flat, need_lo=split)
        cur_c = 3
        segments = []
        for (name, kh, kw, cin, cout, stride, padding, relu), (h, w, oh, ow, pt, pl) in zip(LAYERS, self._geo):
            a_hi, a_lo = ops.im2col_planes(hi, lo, n, h, w, cur_c, kh, kw, stride, pt, pl, oh, ow)
            w_hi, w_lo = st["w"][name]
            pooled = name in POOL_AFTER
            last = name == LAYERS[-1][0]
            out, planes = ops.gemm_planes(a_hi, a_lo, w_hi, w_lo, n * oh * ow, cout, st["b"][name],
                                          "relu" if relu else "none", self.precision, want_f32=True,
                                          want_planes=not pooled and not last)
            segments.append(out)
            cur_c = cout
            if pooled:
                hi, lo, _, _ = ops.maxpool_planes(out, n, oh, ow, cout, 3, 2, need_lo=split)
            elif not last:
                hi, lo = planes
        return ops.cnnvtl_quantise(segments, n, st["keep"])

    def cosine_candidates(self, desc, k=10, db_dtype="fp16"):
        """Loop candidates of BASELINE config 3 ("cnn_vtl descriptors + cosine top-k matching"): the int8 descriptors
        of a sequence (CUDA tensor or ndarray [N, M], as transform returns them) de-quantised to float, every frame
        matched against all the others by cosine similarity on the tensor cores (KeyframeDatabase, fused top-k) ->
        (scores float32 [N, k], indices int64 [N, k]), best first, the frame itself excluded. New capability: the
        reference only fills the dense Hamming matrix (create_distance_matrix.py:27-36)."""
        import torch

        from .matcher import KeyframeDatabase
        d = desc if isinstance(desc, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(desc))
        d = d.cuda().to(torch.float32).contiguous()
        n = d.shape[0]
        k = min(k, max(n - 1, 1))
        db = getattr(self, "_cos_db", None)
        if db is None or db.capacity < n or db.dim != d.shape[1] or db.dtype != db_dtype:
            db = self._cos_db = KeyframeDatabase(d.shape[1], max(n, 1024), "cos", db_dtype)
        db.clear()
        db.append(d)
        s, i = db.topk(d, k + 1)
        # drop the frame itself (cosine 1 with its own stored row; a bit-identical other frame may sort before it)
        self_col = (i == torch.arange(n, device=i.device)[:, None])
        has_self = self_col.any(1, keepdim=True)
        self_col = torch.where(has_self, self_col, torch.nn.functional.one_hot(
            torch.full((n,), k, device=i.device), k + 1).bool())     # not listed: drop the last entry
        keep = ~self_col
        return s[keep].view(n, k), i[keep].view(n, k)

    # images per device pass of transform(): bounds the workspace (~7 MB per 192x240 image); the reference feeds
    # all N images to one session.run and never uses `batch_size` in transform (cnn_vtl.py:130-133)
    DEVICE_CHUNK = 256

    def transform(self, x):
        import torch
        x = np.asarray(x)
        if x.ndim != 4 or x.shape[1] != self._H or x.shape[2] != self._W or x.shape[3] != 3:
            raise ValueError("expected input [N, %d, %d, 3], got %s" % (self._H, self._W, x.shape))
        if x.dtype not in (np.uint8, np.float32, np.float64):
            x = x.astype(np.float64)
        outs = []
        for s in range(0, x.shape[0], self.DEVICE_CHUNK):
            xc = torch.from_numpy(np.ascontiguousarray(x[s:s + self.DEVICE_CHUNK])).cuda()
            outs.append(self._forward_chunk(xc).cpu().numpy())
        return np.concatenate(outs) if outs else np.zeros((0, self.keep_cols.size), dtype=np.int8)
