"""cnn_vtl conv-head oracle: float64 NumPy restatement of src/cnn_vtl/network/cnn_vtl.py:28-128 (graph) and
src/utils/MathUtils.py:1-4. PARITY UNPINNED (TensorFlow 1.x and the AlexNet blob are unavailable); tf.layers.conv2d
semantics ('valid' default, 'same' = TF padding with the extra pixel at the bottom/right, NHWC, HWIO kernels) are
[TF1-doc]."""
import numpy as np

# (name, kh, kw, cin, cout, stride, padding, relu)   cnn_vtl.py:33-93  (ungrouped convs, conv5 linear)
LAYERS = [
    ("conv1", 11, 11, 3, 96, 4, "valid", True),
    ("conv2", 5, 5, 96, 256, 1, "same", True),
    ("conv3", 3, 3, 256, 384, 1, "same", True),
    ("conv4", 3, 3, 384, 384, 1, "same", True),
    ("conv5", 3, 3, 384, 256, 1, "same", False),
]


def compressed_size(value, compression):
    """MathUtils.compressed_size (src/utils/MathUtils.py:1-4)."""
    return int(round(value * ((100 - compression) / 100)))


def make_weights(seed):
    """Synthetic He-scaled HWIO kernels + small biases of the shapes the graph creates (the real blob is an LFS
    pointer)."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, kh, kw, cin, cout, *_ in LAYERS:
        fan_in = kh * kw * cin
        params[name] = (rng.standard_normal((kh, kw, cin, cout)) * np.sqrt(2.0 / fan_in),
                        0.05 * rng.standard_normal(cout))
    return params


def _same_pad(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2, total - total // 2


def conv2d(x, w, b, stride, padding, relu):
    """x [N,H,W,C] float64, w [kh,kw,C,Cout]."""
    n, h, wd, c = x.shape
    kh, kw, _, cout = w.shape
    if padding == "same":
        oh, pt, pb = _same_pad(h, kh, stride)
        ow, pl, pr = _same_pad(wd, kw, stride)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    else:
        oh = (h - kh) // stride + 1
        ow = (wd - kw) // stride + 1
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))  # [N, H', W', C, kh, kw]
    win = win[:, ::stride, ::stride][:, :oh, :ow]
    cols = win.transpose(0, 1, 2, 4, 5, 3).reshape(n * oh * ow, kh * kw * c)   # (kh, kw, c) order = HWIO flatten
    y = cols @ w.reshape(kh * kw * c, cout) + b
    if relu:
        y = np.maximum(y, 0.0)
    return y.reshape(n, oh, ow, cout)


def maxpool(x, k=3, stride=2):
    """tf.layers.max_pooling2d 3x3 / 2, 'valid' (cnn_vtl.py:42-45, 58-61)."""
    win = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(1, 2))[:, ::stride, ::stride]
    return win.max(axis=(-1, -2))


def conv_outputs(x, params):
    outs = []
    h = np.asarray(x, dtype=np.float64)
    for name, kh, kw, cin, cout, stride, padding, relu in LAYERS:
        h = conv2d(h, params[name][0], params[name][1], stride, padding, relu)
        outs.append(h)
        if name in ("conv1", "conv2"):
            h = maxpool(h)
    return outs


def layer_sizes(input_hw):
    h, w = input_hw
    sizes = []
    for name, kh, kw, cin, cout, stride, padding, relu in LAYERS:
        if padding == "same":
            h, w = -(-h // stride), -(-w // stride)
        else:
            h, w = (h - kh) // stride + 1, (w - kw) // stride + 1
        sizes.append(h * w * cout)
        if name in ("conv1", "conv2"):
            h, w = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    return sizes


def make_keep_columns(sizes, compress_factor=99.59, seed=0):
    """Column mask of cnn_vtl.py:119-126: per layer, compressed_size indices drawn WITH replacement; returns the
    sorted unique kept columns (boolean_mask keeps mask order). Seeded here; the reference is unseeded."""
    rng = np.random.default_rng(seed)
    keep = []
    start = 0
    for s in sizes:
        idx = rng.choice(np.arange(start, start + s), size=compressed_size(s, compress_factor))
        keep.append(np.unique(idx))
        start += s
    return np.concatenate(keep).astype(np.int64)


def cast_int8_wrap(v):
    """tf.cast(float64 -> int8) of out-of-range values on x86: convert through int32 (truncate toward zero) and
    keep the low byte [TF1-doc]."""
    t = np.trunc(v)
    t = np.where(np.isfinite(t), t, -2147483648.0)
    return (t.astype(np.int64) & 0xFF).astype(np.uint8).view(np.int8)


def descriptors_from_outputs(outs, keep_cols):
    n = outs[0].shape[0]
    d = np.concatenate([o.reshape(n, -1) for o in outs], axis=1)          # :96-106
    mx = d.max(axis=1, keepdims=True)
    mn = d.min(axis=1, keepdims=True)
    scaled = (d - mn) * (255.0 / (mx - mn))                                 # :109-115
    return cast_int8_wrap(scaled)[:, keep_cols], scaled[:, keep_cols]       # :116, :128


def transform(x, params, keep_cols):
    return descriptors_from_outputs(conv_outputs(x, params), keep_cols)[0]
