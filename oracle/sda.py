"""SDA encoder oracle: float64 forward of SDAV / DA. Follows src/sdav/network/SDAV.py:120-163 (graph),
:188-217 (variables), :293-302 (transform) and src/sdav/network/DenoisingAutoencoderVariant.py:92-101, 116-119.
PARITY UNPINNED (TensorFlow 1.x unavailable); with corruption_level = 0 the corruption mask is all ones
(src/utils/TensorflowWrapper.py:34-38, 148-156), so transform is a plain 5-layer sigmoid MLP."""
import numpy as np

INPUT_SHAPE = [30, 1681]                      # SDAV.py:31
HIDDEN_UNITS = [2500, 2500, 2500, 2500, 2500]  # SDAV.py:32


def sigmoid(z):
    z = np.asarray(z, dtype=np.float64)
    out = np.empty_like(z)
    pos = z >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-z[pos]))
    e = np.exp(z[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def make_weights(dims, seed, scale="normal"):
    """Reference initialisation: tf.random_normal (sigma = 1) weights, zero biases (SDAV.py:189-217).
    scale="xavier" gives trained-like magnitudes (sigma = 1/sqrt(fan_in)) for the precision study."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for k, n in zip(dims[:-1], dims[1:]):
        w = rng.standard_normal((k, n))
        if scale == "xavier":
            w /= np.sqrt(k)
            b = 0.1 * rng.standard_normal(n)
        else:
            b = np.zeros(n)
        ws.append(w)
        bs.append(b)
    return ws, bs


def sda_forward(x, weights, biases):
    """x float64 [B, P, in] (or [rows, in]) -> float64 [B*P, out]: the flat tensor SDAV.transform returns (:163)."""
    h = np.asarray(x, dtype=np.float64)
    h = h.reshape(-1, h.shape[-1])                # flat_batch (TensorflowWrapper.py:13-15)
    for w, b in zip(weights, biases):
        h = sigmoid(h @ w + b)                    # SDAV.py:129,136,143,150,157
    return h
