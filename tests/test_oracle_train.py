"""The training-step oracle (oracle/train.py) pinned two ways (CPU, float64): with `exact_gradient=True` its analytic
gradients of the restated SDAV / DA losses equal central finite differences of the loss itself and torch autograd of
the same loss; its DEFAULT follows TensorFlow's registered gradient of softmax_cross_entropy_with_logits_v2
(backprop = softmax - labels, which is not the derivative when a row's labels do not sum to one) - pinned by a
hand-computed vector and by autograd through a custom op with exactly that backward [TF1-doc]."""
import numpy as np
import pytest

from oracle import train as o_train


def _setup(seed=0, B=3, P=4, dims=(6, 5, 4)):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (B, P, dims[0]))
    Ws = [0.7 * rng.standard_normal((k, n)) for k, n in zip(dims[:-1], dims[1:])]
    bs = [0.1 * rng.standard_normal(n) for n in dims[1:]]
    bds = [0.1 * rng.standard_normal(k) for k in dims[:-1]]
    masks = [o_train.sdav_mask(P, d, 0.3, rng) for d in dims[:-1]]
    return rng, x, Ws, bs, bds, masks


def _fd(f, arr, n_probe, rng, eps=1e-6):
    """central differences of scalar f at n_probe random entries of arr (modified in place and restored)."""
    out = []
    flat = arr.reshape(-1)
    for idx in rng.choice(flat.size, size=min(n_probe, flat.size), replace=False):
        old = flat[idx]
        flat[idx] = old + eps
        fp = f()
        flat[idx] = old - eps
        fm = f()
        flat[idx] = old
        out.append((idx, (fp - fm) / (2 * eps)))
    return out


@pytest.mark.parametrize("layer_i", [0, 1])
def test_sdav_gradients_match_finite_differences(layer_i):
    rng, x, Ws, bs, bds, masks = _setup()
    loss, dW, db, dbd = o_train.sdav_loss_and_grads(x, Ws, bs, bds, layer_i, masks, exact_gradient=True)
    assert np.isclose(loss, o_train.sdav_loss(x, Ws, bs, bds, layer_i, masks))
    f = lambda: o_train.sdav_loss(x, Ws, bs, bds, layer_i, masks)  # noqa: E731
    for l in range(layer_i + 1):
        for idx, g in _fd(f, Ws[l], 12, rng):
            assert abs(g - dW[l].reshape(-1)[idx]) <= 1e-6 * max(1.0, abs(g)), ("W", l, idx)
        for idx, g in _fd(f, bs[l], 4, rng):
            assert abs(g - db[l][idx]) <= 1e-6 * max(1.0, abs(g)), ("b", l, idx)
    for idx, g in _fd(f, bds[layer_i], 4, rng):
        assert abs(g - dbd[idx]) <= 1e-6 * max(1.0, abs(g))
    # decoder biases of lower layers do not enter loss_i
    if layer_i == 1:
        for idx, g in _fd(f, bds[0], 3, rng):
            assert g == 0.0


def test_sdav_quirks_are_reproduced():
    """cs of layer 0 sums over the PATCH axis, cs of deeper layers over the hidden axis (Appendix A.5); labels of
    layer >= 1 are the corrupted inputs; one step on loss_1 also moves layer 0 (no var_list)."""
    rng, x, Ws, bs, bds, masks = _setup(1)
    B, P = x.shape[:2]
    xs, hs, y = o_train.sdav_forward(x, Ws, bs, bds, 0, masks)
    cd, cs, cc = o_train.loss_terms(x.reshape(B * P, -1), y, hs[0], B, P, 0.05, cs_over_patches=True)
    assert np.isclose(cs, np.mean(np.linalg.norm(hs[0].reshape(B, P, -1) - 0.05, ord=1, axis=1)))
    assert np.isclose(cc, np.mean([np.linalg.norm(a - b) for a, b in zip(hs[0].reshape(B, P, -1)[:-1],
                                                                        hs[0].reshape(B, P, -1)[1:])]))
    loss, W2, b2, bd2 = o_train.sdav_train_step(x, Ws, bs, bds, 1, masks)
    assert not np.array_equal(W2[0], Ws[0]) and not np.array_equal(W2[1], Ws[1])
    assert np.array_equal(bd2[0], bds[0]) and not np.array_equal(bd2[1], bds[1])
    # exactly round(P * width * level) zeros per mask
    assert int((masks[0] == 0).sum()) == int(np.round(P * 6 * 0.3))


def test_da_gradients_match_finite_differences():
    rng = np.random.default_rng(2)
    B, P, n_in, hid = 3, 4, 7, 5
    x = rng.uniform(0, 1, (B, P, n_in))
    w0, b0, b1 = 0.7 * rng.standard_normal((n_in, hid)), 0.1 * rng.standard_normal(hid), 0.1 * rng.standard_normal(n_in)
    zm, om = o_train.da_masks(B * P, n_in, 0.3, rng)
    assert int((zm == 0).sum()) == int(B * P * n_in * 0.3) and np.all(om[zm == 1] == 0)
    loss, dW, db0, db1 = o_train.da_loss_and_grads(x, w0, b0, b1, zm, om, exact_gradient=True)
    f = lambda: o_train.da_loss_and_grads(x, w0, b0, b1, zm, om)[0]  # noqa: E731
    for idx, g in _fd(f, w0, 15, rng):
        assert abs(g - dW.reshape(-1)[idx]) <= 1e-6 * max(1.0, abs(g))
    for idx, g in _fd(f, b0, 4, rng):
        assert abs(g - db0[idx]) <= 1e-6 * max(1.0, abs(g))
    for idx, g in _fd(f, b1, 4, rng):
        assert abs(g - db1[idx]) <= 1e-6 * max(1.0, abs(g))


def test_tf_registered_xent_gradient_hand_computed():
    """logits y = [0, ln 3] -> softmax [1/4, 3/4]; labels [2, 1] (sum 3, like a patch row). TensorFlow's op emits
    softmax - labels = [-1.75, -0.25]; the derivative of -sum(L log_softmax) is softmax*sum(L) - L = [-1.25, 1.25]."""
    y = np.array([[0.0, np.log(3.0)]])
    labels = np.array([[2.0, 1.0]])
    h = np.full((1, 2), 0.5)
    dy_tf, dl, _ = o_train.loss_term_grads(labels, y, h, 1, 1, 0.05, 0.0, 0.0, False)
    dy_ex, dl2, _ = o_train.loss_term_grads(labels, y, h, 1, 1, 0.05, 0.0, 0.0, False, exact_gradient=True)
    assert np.allclose(dy_tf, [[-1.75, -0.25]], rtol=0, atol=1e-15)
    assert np.allclose(dy_ex, [[-1.25, 1.25]], rtol=0, atol=1e-15)
    assert np.allclose(dl, [[np.log(4.0), np.log(4.0 / 3.0)]]) and np.array_equal(dl, dl2)
    # labels that sum to one: the two coincide
    one = np.array([[0.25, 0.75]])
    a = o_train.loss_term_grads(one, y, h, 1, 1, 0.05, 0.0, 0.0, False)[0]
    b = o_train.loss_term_grads(one, y, h, 1, 1, 0.05, 0.0, 0.0, False, exact_gradient=True)[0]
    assert np.allclose(a, b, rtol=0, atol=1e-15)


@pytest.mark.parametrize("layer_i", [0, 2])
@pytest.mark.parametrize("exact", [True, False])
def test_sdav_gradients_match_torch_autograd(layer_i, exact):
    """Independent pin of the training oracle: the SDAV loss written with torch ops (float64, CPU) and differentiated
    by autograd - including the gradient that softmax_cross_entropy_with_logits_v2 sends into its labels - gives the
    oracle's analytic gradients. exact=False: the cross-entropy is a custom autograd op whose backward is TensorFlow's
    registered one (grad * (softmax - labels) into the logits, grad * -log_softmax into the labels)."""
    import torch

    class TfXent(torch.autograd.Function):
        @staticmethod
        def forward(ctx, logits, labels):
            ls = torch.log_softmax(logits, dim=1)
            ctx.save_for_backward(ls, labels)
            return -(labels * ls).sum(dim=1)

        @staticmethod
        def backward(ctx, g):
            ls, labels = ctx.saved_tensors
            return g[:, None] * (torch.exp(ls) - labels), g[:, None] * (-ls)

    rng, x, Ws, bs, bds, masks = _setup(5, B=4, P=3, dims=(7, 6, 5, 4))
    loss, dW, db, dbd = o_train.sdav_loss_and_grads(x, Ws, bs, bds, layer_i, masks, exact_gradient=exact)
    B, P, _ = x.shape
    tW = [torch.tensor(w, requires_grad=True) for w in Ws]
    tb = [torch.tensor(b, requires_grad=True) for b in bs]
    tbd = [torch.tensor(b, requires_grad=True) for b in bds]
    cur = torch.tensor(x)
    for l in range(layer_i + 1):
        xc = (cur.reshape(B, P, -1) * torch.tensor(masks[l])[None]).reshape(B * P, -1)
        h = torch.sigmoid(xc @ tW[l] + tb[l])
        cur = h
    y = torch.sigmoid(h @ tW[layer_i].T + tbd[layer_i])
    labels = torch.tensor(x).reshape(B * P, -1) if layer_i == 0 else xc
    cd = torch.mean(-(labels * torch.log_softmax(y, dim=1)).sum(dim=1)) if exact else torch.mean(TfXent.apply(y, labels))
    h3 = h.reshape(B, P, -1)
    cs = torch.mean(torch.abs((h3 if layer_i == 0 else h) - 0.05).sum(dim=1))       # 3-D: patch axis; 2-D: hidden axis
    cc = torch.mean(torch.sqrt(((h3[:-1] - h3[1:]) ** 2).sum(dim=(1, 2))))
    t_loss = cd + 1.0 * cs + 0.2 * cc
    t_loss.backward()
    assert abs(float(t_loss) - loss) <= 1e-12 * abs(loss)
    for l in range(layer_i + 1):
        assert np.allclose(tW[l].grad.numpy(), dW[l], rtol=1e-10, atol=1e-13)
        assert np.allclose(tb[l].grad.numpy(), db[l], rtol=1e-10, atol=1e-13)
    assert np.allclose(tbd[layer_i].grad.numpy(), dbd, rtol=1e-10, atol=1e-13)
