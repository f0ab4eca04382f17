"""Training-step oracle (TEST INFRASTRUCTURE ONLY): float64 NumPy restatement of the reference's losses and of one
plain-SGD step, for the two training graphs of the reference.

  SDAV (src/sdav/network/SDAV.py): model :120-163, loss :171-186, optimiser :223-226, fit :242-288.
      layer l:  x_l = corrupt(h_{l-1})  (x_0 = corrupt(input)),  h_l = sigmoid(x_l W_l + b_l),
                y_l = sigmoid(h_l W_l^T + bd_l)                                  (tied decoder weights, :192, :198 ...)
      loss_i = cd_i + sparse_penalty * cs_i + consecutive_penalty * cc_i
        cd_i = mean_rows softmax_cross_entropy_with_logits_v2(labels = L_i, logits = y_i)
               L_0 = the CLEAN input (x0.to_tf() is the placeholder, :130), L_i = the CORRUPTED x_i for i >= 1 (:137 ...);
               the _v2 op back-propagates into its labels, so for i >= 1 the label gradient reaches the lower layers.
        cs_i = mean(norm(h_i - sparse_level, axis=1, ord=1)): h_0 is 3-D [B, P, H] -> the norm runs over the PATCH
               axis; h_i (i >= 1) is 2-D [B*P, H] -> over the hidden axis (Appendix A.5 quirk, reproduced).
        cc_i = mean_b || h_i[b] - h_i[b+1] ||_F over the B-1 consecutive frame pairs.
      `optimizer.minimize(loss_i)` has no var_list: one step updates every trainable variable loss_i depends on -
      W_l, b_l for l <= i and bd_i (:226, Appendix A.5).
      corrupt (src/utils/TensorflowWrapper.py:34-38, 148-156): one [P, in_l] 0/1 mask with exactly
      round(P * in_l * level) zeros, shared by the whole batch, redrawn on every session.run.

  DA (src/sdav/network/DenoisingAutoencoderVariant.py): model :103-119, loss :121-148, corruption :182-202.
      x~ = zeros_mask * x + ones_mask (masks drawn ONCE at graph construction, shape [B*P, in]); labels = the clean
      batch; cs over the hidden axis; variables w0, b0, b1.

PARITY UNPINNED: TensorFlow 1.x is not installable here and the reference has no training test. What pins this file
is (i) the gradient check of tests/test_oracle_train.py (analytic gradients == central finite differences of the
restated loss) and (ii) the citations above.
"""
import numpy as np


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def log_softmax(y):
    m = y.max(axis=1, keepdims=True)
    return y - (m + np.log(np.exp(y - m).sum(axis=1, keepdims=True)))


def sdav_mask(P, width, level, rng):
    """random_mask (TensorflowWrapper.py:148-156): exactly round(P*width*level) zeros, shuffled."""
    n = P * width
    n_zeros = int(np.round(n * level))
    m = np.concatenate([np.ones(n - n_zeros), np.zeros(n_zeros)])
    rng.shuffle(m)
    return m.reshape(P, width)


def da_masks(rows, width, level, rng):
    """_corrupt_tensor (DenoisingAutoencoderVariant.py:182-202): zeros_mask has int(n*level) zeros; half of the
    zeroed positions (Bernoulli 0.5) are set to one by ones_mask."""
    n = rows * width
    zeros_mask = np.ones(n)
    zeros_mask[:int(n * level)] = 0
    rng.shuffle(zeros_mask)
    ones_mask = ((1 - zeros_mask).astype(int) & (rng.random(n) < 0.5).astype(int)).astype(np.float64)
    return zeros_mask.reshape(rows, width), ones_mask.reshape(rows, width)


# ------------------------------------------------------------------------------------------------ loss terms
def loss_terms(labels, y, h, B, P, sparse_level, cs_over_patches):
    """(cd, cs, cc) for flattened labels / y [B*P, in] and h [B*P, hid]."""
    R = B * P
    cd = float(np.mean(-(labels * log_softmax(y)).sum(axis=1)))
    h3 = h.reshape(B, P, -1)
    if cs_over_patches:
        cs = float(np.mean(np.abs(h3 - sparse_level).sum(axis=1)))        # [B, hid] norms
    else:
        cs = float(np.mean(np.abs(h - sparse_level).sum(axis=1)))         # [R] norms
    d = h3[:-1] - h3[1:]
    cc = float(np.mean(np.sqrt((d * d).sum(axis=(1, 2)))))
    assert R == labels.shape[0]
    return cd, cs, cc


def loss_term_grads(labels, y, h, B, P, sparse_level, sparse_penalty, consecutive_penalty, cs_over_patches,
                    exact_gradient=False):
    """d loss / d y, d loss / d labels, d loss / d h (the direct cs + cc part) for loss = cd + sp*cs + cp*cc.

    d cd / d y: TensorFlow's `softmax_cross_entropy_with_logits_v2` (SDAV.py:172, DenoisingAutoencoderVariant.py:124)
    does NOT differentiate its loss exactly when the labels of a row do not sum to one: the op emits
    backprop = softmax(logits) - labels (tensorflow/core/kernels/xent_op.h) and the registered gradient only scales it
    by the incoming gradient (python/ops/nn_grad.py: _SoftmaxCrossEntropyWithLogitsGrad) [TF1-doc]. The labels here are
    image patches / activations that sum to hundreds per row, so `optimizer.minimize` in the reference follows
    (softmax - labels) / R - the default here - not the mathematical derivative (softmax * sum(labels) - labels) / R
    (`exact_gradient=True`; that one equals finite differences of the loss). The label gradient -log_softmax / R and
    the loss value are the same in both."""
    R = B * P
    ls = log_softmax(y)
    sm = np.exp(ls)
    dy = (sm * (labels.sum(axis=1, keepdims=True) if exact_gradient else 1.0) - labels) / R
    dlabels = -ls / R
    hid = h.shape[1]
    count = B * hid if cs_over_patches else R
    dh = sparse_penalty * np.sign(h - sparse_level) / count
    h3 = h.reshape(B, P, hid)
    d = h3[:-1] - h3[1:]
    nrm = np.sqrt((d * d).sum(axis=(1, 2)))
    g = d / nrm[:, None, None] / (B - 1)
    dcc = np.zeros_like(h3)
    dcc[:-1] += g
    dcc[1:] -= g
    dh = dh + consecutive_penalty * dcc.reshape(R, hid)
    return dy, dlabels, dh


# ------------------------------------------------------------------------------------------------ SDAV
def sdav_forward(x, Ws, bs, bds, layer_i, masks):
    """x [B, P, in] -> per-layer corrupted inputs xs[l] [B*P, in_l], hiddens hs[l], and the decoder output y_i."""
    B, P, _ = x.shape
    cur = x.reshape(B * P, -1)
    xs, hs = [], []
    for l in range(layer_i + 1):
        xc = (cur.reshape(B, P, -1) * masks[l][None]).reshape(B * P, -1)
        h = sigmoid(xc @ Ws[l] + bs[l])
        xs.append(xc)
        hs.append(h)
        cur = h
    y = sigmoid(hs[layer_i] @ Ws[layer_i].T + bds[layer_i])
    return xs, hs, y


def sdav_loss(x, Ws, bs, bds, layer_i, masks, sparse_level=0.05, sparse_penalty=1.0, consecutive_penalty=0.2):
    B, P, _ = x.shape
    xs, hs, y = sdav_forward(x, Ws, bs, bds, layer_i, masks)
    labels = x.reshape(B * P, -1) if layer_i == 0 else xs[layer_i]
    cd, cs, cc = loss_terms(labels, y, hs[layer_i], B, P, sparse_level, cs_over_patches=(layer_i == 0))
    return cd + sparse_penalty * cs + consecutive_penalty * cc


def sdav_loss_and_grads(x, Ws, bs, bds, layer_i, masks, sparse_level=0.05, sparse_penalty=1.0,
                        consecutive_penalty=0.2, exact_gradient=False):
    """loss_i and its gradients: (loss, dW[0..i], db[0..i], dbd_i)."""
    B, P, _ = x.shape
    xs, hs, y = sdav_forward(x, Ws, bs, bds, layer_i, masks)
    i = layer_i
    labels = x.reshape(B * P, -1) if i == 0 else xs[i]
    cd, cs, cc = loss_terms(labels, y, hs[i], B, P, sparse_level, cs_over_patches=(i == 0))
    loss = cd + sparse_penalty * cs + consecutive_penalty * cc
    dy, dlabels, dh = loss_term_grads(labels, y, hs[i], B, P, sparse_level, sparse_penalty, consecutive_penalty,
                                      cs_over_patches=(i == 0), exact_gradient=exact_gradient)
    dzy = dy * y * (1 - y)
    dbd = dzy.sum(axis=0)
    dW = [None] * (i + 1)
    db = [None] * (i + 1)
    dW[i] = dzy.T @ hs[i]                         # decoder use of W_i (as W_i^T): d/dW_i = (h^T dzy)^T
    dh = dh + dzy @ Ws[i]
    dx_extra = dlabels if i > 0 else None         # label gradient reaches x_i (and below) only for i >= 1
    for l in range(i, -1, -1):
        dzh = dh * hs[l] * (1 - hs[l])
        dW[l] = (dW[l] if dW[l] is not None else 0) + xs[l].T @ dzh
        db[l] = dzh.sum(axis=0)
        if l == 0:
            break
        dx = dzh @ Ws[l].T
        if dx_extra is not None:
            dx = dx + dx_extra
            dx_extra = None
        dh = (dx.reshape(B, P, -1) * masks[l][None]).reshape(B * P, -1)      # x_l = h_{l-1} * mask_l
    return loss, dW, db, dbd


def sdav_train_step(x, Ws, bs, bds, layer_i, masks, lr=0.1, **kw):
    """One GradientDescentOptimizer step on loss_i; returns (loss before the step, new Ws, bs, bds)."""
    loss, dW, db, dbd = sdav_loss_and_grads(x, Ws, bs, bds, layer_i, masks, **kw)
    Ws, bs, bds = [w.copy() for w in Ws], [b.copy() for b in bs], [b.copy() for b in bds]
    for l in range(layer_i + 1):
        Ws[l] -= lr * dW[l]
        bs[l] -= lr * db[l]
    bds[layer_i] -= lr * dbd
    return loss, Ws, bs, bds


# ------------------------------------------------------------------------------------------------ DA
def da_forward(x, w0, b0, b1, zeros_mask, ones_mask):
    B, P, _ = x.shape
    xc = zeros_mask * x.reshape(B * P, -1) + ones_mask
    h = sigmoid(xc @ w0 + b0)
    y = sigmoid(h @ w0.T + b1)
    return xc, h, y


def da_loss_and_grads(x, w0, b0, b1, zeros_mask, ones_mask, sparse_level=0.05, sparse_penalty=1.0,
                      consecutive_penalty=0.2, exact_gradient=False):
    B, P, _ = x.shape
    xc, h, y = da_forward(x, w0, b0, b1, zeros_mask, ones_mask)
    labels = x.reshape(B * P, -1)
    cd, cs, cc = loss_terms(labels, y, h, B, P, sparse_level, cs_over_patches=False)
    loss = cd + sparse_penalty * cs + consecutive_penalty * cc
    dy, _, dh = loss_term_grads(labels, y, h, B, P, sparse_level, sparse_penalty, consecutive_penalty, False,
                                exact_gradient=exact_gradient)
    dzy = dy * y * (1 - y)
    dh = dh + dzy @ w0
    dzh = dh * h * (1 - h)
    dW = dzy.T @ h + xc.T @ dzh
    return loss, dW, dzh.sum(axis=0), dzy.sum(axis=0)


def da_train_step(x, w0, b0, b1, zeros_mask, ones_mask, lr=0.1, **kw):
    loss, dW, db0, db1 = da_loss_and_grads(x, w0, b0, b1, zeros_mask, ones_mask, **kw)
    return loss, w0 - lr * dW, b0 - lr * db0, b1 - lr * db1
