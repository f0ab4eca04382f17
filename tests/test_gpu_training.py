"""One SGD step of the denoising-autoencoder training graphs on the B200 (dlc_train_* kernels + tcgen05 GEMMs through
the C ABI) against the float64 oracle (oracle/train.py, itself pinned by finite differences). `-m gpu`."""
import numpy as np
import pytest
import torch

from oracle import train as o_train

pytestmark = pytest.mark.gpu


def _nerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _setup(dims, B, P, seed, scale):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (B, P, dims[0]))
    Ws = [scale * rng.standard_normal((k, n)) for k, n in zip(dims[:-1], dims[1:])]
    bs = [0.1 * rng.standard_normal(n) for n in dims[1:]]
    bds = [0.1 * rng.standard_normal(k) for k in dims[:-1]]
    return rng, x, Ws, bs, bds


# exact = False: TensorFlow's registered cross-entropy gradient (softmax - labels), the reference-faithful default;
# exact = True: the mathematical derivative of the loss (see dlc_train_xent_grad)
@pytest.mark.parametrize("dims,B,P,top,exact", [((70, 96, 40), 4, 5, 0, False), ((70, 96, 40), 4, 5, 1, False),
                                                ((70, 96, 40), 4, 5, 1, True), ((300, 130, 64, 33), 3, 30, 2, False),
                                                ((300, 130, 64, 33), 3, 30, 2, True), ((1681, 2500), 10, 30, 0, False)])
def test_sdav_train_step_vs_oracle(cuda, dims, B, P, top, exact):
    from deeploopcloser_b200.training import DaeStackTrainer
    rng, x, Ws, bs, bds = _setup(dims, B, P, 7 + top, 0.3 if dims[0] < 1000 else 0.05)
    masks = [o_train.sdav_mask(P, d, 0.3, rng) for d in dims[:top + 1]]
    tr = DaeStackTrainer(dims, patches=P, exact_gradient=exact)
    tr.set_weights(Ws, bs, bds)
    xd = torch.from_numpy(x).float().cuda()
    md = [torch.from_numpy(m).float().cuda() for m in masks]
    loss = float(tr.step(xd, top, md).item())
    want_loss, dW, db, dbd = o_train.sdav_loss_and_grads(x, Ws, bs, bds, top, masks, exact_gradient=exact)
    g = tr.last_grads
    print("loss %.9f vs %.9f" % (loss, want_loss))
    assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)
    for l in range(top + 1):
        e_w, e_b = _nerr(g["dW"][l].cpu().numpy(), dW[l]), _nerr(g["db"][l].cpu().numpy(), db[l])
        print("layer %d: dW normwise err %.2e, db %.2e" % (l, e_w, e_b))
        assert e_w <= 1e-4 and e_b <= 1e-4
    assert _nerr(g["dbd"].cpu().numpy(), dbd) <= 1e-4
    # the update itself: every layer <= top moved by lr * grad, the decoder bias only at `top`
    _, W2, b2, bd2 = o_train.sdav_train_step(x, Ws, bs, bds, top, masks, lr=0.1, exact_gradient=exact)
    gW, gb, gbd = tr.get_weights()
    for l in range(len(dims) - 1):
        assert _nerr(gW[l], W2[l]) <= 1e-6 and _nerr(gb[l], b2[l]) <= 1e-5
        assert np.allclose(gbd[l], bd2[l], rtol=0, atol=1e-6)
    assert tr.global_step == 1


def test_da_train_step_vs_oracle(cuda):
    from deeploopcloser_b200.training import DaeStackTrainer
    dims, B, P = (120, 200), 5, 6
    rng, x, Ws, bs, bds = _setup(dims, B, P, 3, 0.3)
    zm, om = o_train.da_masks(B * P, dims[0], 0.3, rng)
    tr = DaeStackTrainer(dims, patches=P)
    tr.set_weights(Ws, bs, bds)
    loss = float(tr.step(torch.from_numpy(x).float().cuda(), 0, [torch.from_numpy(zm).float().cuda()],
                         [torch.from_numpy(om).float().cuda()], mask_rows=B * P, da_mode=True).item())
    want_loss, dW, db0, db1 = o_train.da_loss_and_grads(x, Ws[0], bs[0], bds[0], zm, om)
    g = tr.last_grads
    assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)
    assert _nerr(g["dW"][0].cpu().numpy(), dW) <= 1e-4
    assert _nerr(g["db"][0].cpu().numpy(), db0) <= 1e-4 and _nerr(g["dbd"].cpu().numpy(), db1) <= 1e-4


def test_training_reduces_the_loss(cuda):
    """A few SGD steps on a fixed batch and fixed masks lower loss_0 monotonically (sanity of the whole chain)."""
    from deeploopcloser_b200.training import DaeStackTrainer
    dims, B, P = (64, 48), 6, 4
    rng, x, Ws, bs, bds = _setup(dims, B, P, 11, 0.2)
    # exact_gradient: descent on the loss itself (TensorFlow's registered cross-entropy gradient, the default, is not
    # the derivative of this loss on un-normalised labels, so monotone decrease is only guaranteed for the exact one)
    tr = DaeStackTrainer(dims, patches=P, learning_rate=0.05, exact_gradient=True)
    tr.set_weights(Ws, bs, bds)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0)
    masks = tr.sdav_masks(0, 0.3, gen)
    assert int((masks[0] == 0).sum().item()) == int(round(P * dims[0] * 0.3))
    xd = torch.from_numpy(x).float().cuda()
    losses = [float(tr.step(xd, 0, masks).item()) for _ in range(8)]
    print("losses", losses)
    assert all(b < a for a, b in zip(losses, losses[1:]))


def test_sdav_and_da_fit_api(cuda, tmp_path):
    """The reference's training entry points (train.py -> SDAV.fit_dataset, train-sdav.py -> SDA.fit -> DA.fit_dataset)
    run on the B200: the loss falls, weights move for every layer, and the TensorFlow-format checkpoint written after
    each layer restores into a fresh instance that encodes identically."""
    import logging
    from src.sdav.network.SDAV import SDAV
    from src.sdav.network.StackedDenoisingAutoencoderVariants import SDA
    rng = np.random.default_rng(0)
    net = SDAV(verbosity=logging.ERROR, train_path=str(tmp_path / "sdav"), seed=1, input_shape=[6, 49],
               hidden_units=[40, 32])
    net.epochs, net.default_batch_size, net.learning_rate = 3, 4, 0.05
    net.set_weights([0.2 * w for w in net._weights], net._biases)
    w_before = [w.copy() for w in net._weights]
    frames = rng.uniform(0, 1, (9, 6, 49))                 # 2 batches of 4 + a single trailing frame (ignored)
    first = net.fit(frames[:4], seed=3)
    assert len(first) == 2 and all(np.isfinite(first))
    net.fit_dataset(frames, seed=4)
    assert all(not np.array_equal(a, b) for a, b in zip(w_before, net._weights))
    assert net.global_step == 2 * 3 + 2 * 2 * 3            # fit: 2 layers x 3 epochs; fit_dataset: 2 layers x 2 batches x 3
    # a fresh instance restores the latest checkpoint of the directory by itself (SDAV._load_or_init_session)
    again = SDAV(verbosity=logging.ERROR, train_path=str(tmp_path / "sdav"), input_shape=[6, 49], hidden_units=[40, 32])
    assert again.global_step == net.global_step
    assert all(np.array_equal(a, b) for a, b in zip(again._weights, net._weights))
    assert np.array_equal(again.transform(frames[:2]), net.transform(frames[:2]))

    sda = SDA([6, 49], [40, 32], batch_size=4, epochs=4, learning_rate=0.05, seed=2)
    for layer in sda._layers:
        layer.set_weights(0.2 * layer._w0, layer._b0)
    losses = sda.fit_dataset(list(frames))
    assert len(losses) == 2 and all(np.isfinite(losses))
    l0 = sda._layers[0]
    first_loss = l0.fit_dataset(list(frames[:4]))           # continues from the trained weights, same fixed masks
    assert first_loss < 1e9 and l0.global_step == 2 * 4 + 4


@pytest.mark.parametrize("dims,top", [((70, 96, 40), 1), ((300, 130, 64, 33), 2)])
def test_graphed_step_equals_eager(cuda, dims, top):
    """DaeStackTrainer.graphed_step: the SGD step replayed as a CUDA graph gives the same losses (to the last bits of
    their atomically summed float64) and bit-identical weights as the eager step over several steps with fresh masks;
    capturing performs no update."""
    from deeploopcloser_b200.training import DaeStackTrainer
    B, P = 4, 5
    rng, x, Ws, bs, bds = _setup(dims, B, P, 11, 0.3)
    xd = torch.from_numpy(x).float().cuda()
    a, b = DaeStackTrainer(dims, patches=P), DaeStackTrainer(dims, patches=P)
    a.set_weights(Ws, bs, bds)
    b.set_weights(Ws, bs, bds)
    mask_sets = [[torch.from_numpy(o_train.sdav_mask(P, d, 0.3, rng)).float().cuda() for d in dims[:top + 1]]
                 for _ in range(3)]
    g = b.graphed_step(xd, top, mask_sets[0])
    assert b.global_step == 0
    for w0, w1 in zip(a.get_weights()[0], b.get_weights()[0]):
        assert np.array_equal(w0, w1)                      # nothing moved during warm-up / capture
    for masks in mask_sets:
        la = float(a.step(xd, top, masks).item())
        lb = float(g(xd, masks).item())
        assert abs(la - lb) <= 1e-13 * abs(la)          # the loss is summed with float64 atomics: last-bit order effects
    assert a.global_step == b.global_step == 3
    for wa, wb in zip(a.get_weights(), b.get_weights()):
        for u, v in zip(wa, wb):
            assert np.array_equal(u, v)
