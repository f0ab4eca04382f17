"""Training step of the reference's denoising autoencoders on the B200.

`DaeStackTrainer` holds float64 master weights on the device and runs one plain-SGD step of
  * SDAV's `train_steps[i]` (src/sdav/network/SDAV.py:120-186, 223-226): loss_i through layers 0..i, updating W_l, b_l
    for every l <= i and the decoder bias of layer i (the reference's `minimize` has no var_list), and
  * DA's `train_step` (src/sdav/network/DenoisingAutoencoderVariant.py:103-158): one layer, salt-and-mask corruption
    with masks fixed at construction.
Every contraction is a tcgen05 GEMM through the C ABI (`dlc_gemm_planes`, fp16 hi/lo split); the loss terms, their
gradients, the bias sums, the transposes and the SGD update are the `dlc_train_*` kernels. This module only sequences
those calls - as the reference's Python sequences TensorFlow ops - and owns the device buffers (torch tensors).
There is no CPU path."""
import gc

import numpy as np
import torch

from . import _lib, ops
from ._cuda import ptr, stream_ptr


def _pad(n):
    """Padded width of a layer: multiple of 256 above 256 (so a 256-wide accumulator tile divides it), else of 64."""
    return (n + 255) // 256 * 256 if n > 256 else (n + 63) // 64 * 64


class _Layer:
    def __init__(self, n_in, n_hid, device):
        self.n_in, self.n_hid = n_in, n_hid
        self.in_pad, self.hid_pad = _pad(n_in), _pad(n_hid)
        z = lambda *shape, dt=torch.float16: torch.zeros(shape, dtype=dt, device=device)  # noqa: E731
        self.W = z(n_in, n_hid, dt=torch.float64)
        self.b = z(n_hid, dt=torch.float64)
        self.bd = z(n_in, dt=torch.float64)
        # operand planes of the weights in both orientations (rows beyond the valid ones stay zero)
        self.wt_hi, self.wt_lo = z(self.hid_pad, self.in_pad), z(self.hid_pad, self.in_pad)   # [hid, in]: x W, dzy W
        self.w_hi, self.w_lo = z(self.in_pad, self.hid_pad), z(self.in_pad, self.hid_pad)     # [in, hid]: h W^T, dzh W^T
        self.b32 = z(self.hid_pad, dt=torch.float32)
        self.bd32 = z(self.in_pad, dt=torch.float32)

    def repack(self):
        _lib.call("dlc_pack_weight_planes", ptr(self.W), _lib.F64, self.n_in, self.n_hid, self.hid_pad, ptr(self.wt_hi),
                  ptr(self.wt_lo), self.in_pad, stream_ptr())
        _lib.call("dlc_split_planes", ptr(self.W), _lib.F64, self.n_in, self.n_hid, self.n_hid, 1, 1, ptr(self.w_hi),
                  ptr(self.w_lo), self.hid_pad, stream_ptr())
        self.b32[:self.n_hid].copy_(self.b)
        self.bd32[:self.n_in].copy_(self.bd)


class DaeStackTrainer:
    def __init__(self, dims, patches=30, sparse_level=0.05, sparse_penalty=1.0, consecutive_penalty=0.2,
                 learning_rate=0.1, device="cuda", exact_gradient=False):
        from . import _cuda
        _cuda.require_cuda()
        self.dims = [int(d) for d in dims]
        self.P = int(patches)
        self.sparse_level = float(sparse_level)
        self.sparse_penalty = float(sparse_penalty)
        self.consecutive_penalty = float(consecutive_penalty)
        self.lr = float(learning_rate)
        # False (reference-faithful): the cross-entropy term back-propagates TensorFlow's registered gradient
        # (softmax - labels) / R, what optimizer.minimize follows in the reference; True: the mathematical derivative
        # of the loss (they differ because patch rows do not sum to one, see dlc_train_xent_grad)
        self.exact_gradient = bool(exact_gradient)
        self.device = torch.device(device)
        self.layers = [_Layer(k, n, self.device) for k, n in zip(self.dims[:-1], self.dims[1:])]
        self.loss = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.global_step = 0

    # ---- weights
    def set_weights(self, Ws, bs, bds=None):
        for l, L in enumerate(self.layers):
            L.W.copy_(torch.from_numpy(np.ascontiguousarray(Ws[l], dtype=np.float64)))
            L.b.copy_(torch.from_numpy(np.ascontiguousarray(bs[l], dtype=np.float64)))
            if bds is not None:
                L.bd.copy_(torch.from_numpy(np.ascontiguousarray(bds[l], dtype=np.float64)))
            else:
                L.bd.zero_()
            L.repack()

    def get_weights(self):
        return ([L.W.cpu().numpy() for L in self.layers], [L.b.cpu().numpy() for L in self.layers],
                [L.bd.cpu().numpy() for L in self.layers])

    # ---- one GEMM: act(A B^T + bias) -> float32 [m, n] (+ planes [m, pad(n)])
    @staticmethod
    def _gemm(a_hi, a_lo, b_hi, b_lo, m, n, bias, act, want_planes):
        return ops.gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, bias, act, "fp16x2", want_f32=True, want_planes=want_planes)

    def _planes(self, rows, ld):
        return (torch.empty((rows, ld), dtype=torch.float16, device=self.device),
                torch.empty((rows, ld), dtype=torch.float16, device=self.device))

    # ---- the step
    def step(self, x, top, keep_masks, add_masks=None, mask_rows=None, da_mode=False, apply_update=True):
        """One SGD step on the loss of layer `top` for a batch x [B, P, dims[0]] (float32 CUDA).
        keep_masks[l] / add_masks[l]: float32 CUDA masks [mask_rows, dims[l]] for l <= top (None = no corruption).
        da_mode: DenoisingAutoencoderVariant semantics (labels = the clean batch, sparsity over the hidden axis);
        otherwise SDAV semantics (layer 0: clean labels, sparsity over the patch axis; layer >= 1: the corrupted
        input is the label and receives the label gradient). Returns the loss (float64 device tensor, pre-step)."""
        B, P = int(x.shape[0]), int(x.shape[1])
        if P != self.P or x.shape[2] != self.dims[0]:
            raise ValueError("expected a batch [B, %d, %d]" % (self.P, self.dims[0]))
        if B < 2:
            raise ValueError("the consecutive-frame term needs at least two frames per batch")
        R = B * P
        mask_rows = mask_rows or P
        st = stream_ptr()
        x2 = x.reshape(R, self.dims[0]).to(torch.float32).contiguous()
        self.loss.zero_()

        # ---- forward through layers 0..top
        xs, hs, x_planes, h_planes = [], [], [], []
        cur = x2
        for l in range(top + 1):
            L = self.layers[l]
            xc = torch.empty((R, L.n_in), dtype=torch.float32, device=self.device)
            xh, xl = self._planes(R, L.in_pad)
            keep = keep_masks[l] if keep_masks is not None else None
            add = add_masks[l] if add_masks is not None else None
            _lib.call("dlc_train_corrupt", ptr(cur), ptr(keep), ptr(add), R, L.n_in, mask_rows, ptr(xc), ptr(xh), ptr(xl),
                      L.in_pad, st)
            h, hp = self._gemm(xh, xl, L.wt_hi, L.wt_lo, R, L.n_hid, L.b32, "sigmoid", True)
            xs.append(xc)
            hs.append(h)
            x_planes.append((xh, xl))
            h_planes.append(hp)
            cur = h
        Lt = self.layers[top]
        y, _ = self._gemm(h_planes[top][0], h_planes[top][1], Lt.w_hi, Lt.w_lo, R, Lt.n_in, Lt.bd32, "sigmoid", False)

        # ---- loss of the top layer and its gradients
        label_grad = (not da_mode) and top > 0
        labels = x2 if (da_mode or top == 0) else xs[top]
        dzy = torch.empty((R, Lt.n_in), dtype=torch.float32, device=self.device)
        dzy_h, dzy_l = self._planes(R, Lt.in_pad)
        dlabel = torch.empty_like(dzy) if label_grad else None
        _lib.call("dlc_train_xent_grad", ptr(y), ptr(labels), R, Lt.n_in, ptr(dzy), ptr(dzy_h), ptr(dzy_l), Lt.in_pad,
                  ptr(dlabel), ptr(self.loss), int(self.exact_gradient), st)
        dbd = torch.empty(Lt.n_in, dtype=torch.float64, device=self.device)
        _lib.call("dlc_train_colsum", ptr(dzy), R, Lt.n_in, ptr(dbd), st)
        dh_rec, _ = self._gemm(dzy_h, dzy_l, Lt.wt_hi, Lt.wt_lo, R, Lt.n_hid, None, "none", False)   # dzy W

        cs_over_patches = (not da_mode) and top == 0
        count = B * Lt.n_hid if cs_over_patches else R
        norms = torch.empty(max(B - 1, 1), dtype=torch.float64, device=self.device)
        grads = []
        dh_up = None
        for l in range(top, -1, -1):
            L = self.layers[l]
            is_top = l == top
            dzh = torch.empty((R, L.n_hid), dtype=torch.float32, device=self.device)
            dzh_h, dzh_l = self._planes(R, L.hid_pad)
            _lib.call("dlc_train_hidden_grad", ptr(hs[l]), ptr(dh_rec if is_top else None), ptr(dh_up), B, P, L.n_hid,
                      self.sparse_level, self.sparse_penalty / count if is_top else 0.0,
                      self.consecutive_penalty / (B - 1) if is_top else 0.0, ptr(norms), ptr(dzh), ptr(dzh_h),
                      ptr(dzh_l), L.hid_pad, ptr(self.loss) if is_top else None, st)
            db = torch.empty(L.n_hid, dtype=torch.float64, device=self.device)
            _lib.call("dlc_train_colsum", ptr(dzh), R, L.n_hid, ptr(db), st)
            # weight gradient: x~^T dzh (+ dzy^T h for the tied decoder of the top layer) as ONE contraction over rows
            rp = (R + 63) // 64 * 64
            segs = 2 if is_top else 1
            a_h = torch.zeros((L.in_pad, segs * rp), dtype=torch.float16, device=self.device)
            a_l = torch.zeros_like(a_h)
            b_h = torch.zeros((L.hid_pad, segs * rp), dtype=torch.float16, device=self.device)
            b_l = torch.zeros_like(b_h)
            _lib.call("dlc_train_transpose_planes", ptr(xs[l]), R, L.n_in, ptr(a_h), ptr(a_l), segs * rp, 0, rp, st)
            _lib.call("dlc_train_transpose_planes", ptr(dzh), R, L.n_hid, ptr(b_h), ptr(b_l), segs * rp, 0, rp, st)
            if is_top:
                _lib.call("dlc_train_transpose_planes", ptr(dzy), R, L.n_in, ptr(a_h), ptr(a_l), segs * rp, rp, rp, st)
                _lib.call("dlc_train_transpose_planes", ptr(hs[l]), R, L.n_hid, ptr(b_h), ptr(b_l), segs * rp, rp, rp, st)
            dW, _ = self._gemm(a_h, a_l, b_h, b_l, L.n_in, L.n_hid, None, "none", False)
            grads.append((l, dW, db))
            if l == 0:
                break
            dx, _ = self._gemm(dzh_h, dzh_l, L.w_hi, L.w_lo, R, L.n_in, None, "none", False)        # dzh W^T
            keep = keep_masks[l] if keep_masks is not None else None
            dh_up = torch.empty((R, L.n_in), dtype=torch.float32, device=self.device)
            _lib.call("dlc_train_mask_grad", ptr(dx), ptr(dlabel if (is_top and label_grad) else None), ptr(keep), R,
                      L.n_in, mask_rows, ptr(dh_up), st)

        loss = self.loss.clone()
        if apply_update:
            for l, dW, db in grads:
                L = self.layers[l]
                _lib.call("dlc_train_sgd", ptr(L.W), ptr(dW), _lib.F32, L.W.numel(), self.lr, st)
                _lib.call("dlc_train_sgd", ptr(L.b), ptr(db), _lib.F64, L.b.numel(), self.lr, st)
            _lib.call("dlc_train_sgd", ptr(Lt.bd), ptr(dbd), _lib.F64, Lt.bd.numel(), self.lr, st)
            for l, _, _ in grads:
                self.layers[l].repack()
            self.global_step += 1
        self.last_grads = {"dW": {l: dW for l, dW, _ in grads}, "db": {l: db for l, _, db in grads}, "dbd": dbd}
        return loss

    def graphed_step(self, x, top, keep_masks, add_masks=None, mask_rows=None, da_mode=False):
        """The same step captured once in a CUDA graph (see GraphedStep): for the fit loops, which repeat one step
        shape `epochs` times per batch and are launch-bound at the reference's batch of 10 frames."""
        return GraphedStep(self, x, top, keep_masks, add_masks, mask_rows, da_mode)

    def cached_graphed_step(self, x, top, keep_masks, add_masks=None, mask_rows=None, da_mode=False):
        """graphed_step, captured once per (batch shape, loss layer, mask layout) and reused across the batches of a
        fit loop (a capture costs two warm-up steps, a graph and a private memory pool)."""
        cache = self.__dict__.setdefault("_graph_cache", {})
        key = (tuple(x.shape), int(top), mask_rows, bool(da_mode), add_masks is not None)
        g = cache.get(key)
        if g is None:
            g = cache[key] = self.graphed_step(x, top, keep_masks, add_masks, mask_rows, da_mode)
        return g

    # ---- mask generators (device-side draws; the reference's are unseeded NumPy / TensorFlow shuffles)
    def sdav_masks(self, top, level, generator=None):
        """One [P, in_l] masking-noise mask per layer l <= top with exactly round(P * in_l * level) zeros
        (src/utils/TensorflowWrapper.py:34-38, 148-156)."""
        out = []
        for l in range(top + 1):
            n = self.P * self.dims[l]
            n_zero = int(round(n * level))
            perm = torch.randperm(n, device=self.device, generator=generator)
            m = torch.ones(n, dtype=torch.float32, device=self.device)
            m[perm[:n_zero]] = 0.0
            out.append(m.reshape(self.P, self.dims[l]))
        return out

    def da_masks(self, rows, level, generator=None):
        """(zeros_mask, ones_mask) [rows, dims[0]] of DA._corrupt_tensor (DenoisingAutoencoderVariant.py:182-202)."""
        n = rows * self.dims[0]
        perm = torch.randperm(n, device=self.device, generator=generator)
        zm = torch.ones(n, dtype=torch.float32, device=self.device)
        zm[perm[:int(n * level)]] = 0.0
        coin = torch.rand(n, device=self.device, generator=generator) < 0.5
        om = ((zm == 0) & coin).to(torch.float32)
        return zm.reshape(rows, self.dims[0]), om.reshape(rows, self.dims[0])


class GraphedStep:
    """One `DaeStackTrainer.step` (fixed batch shape, loss layer and mask layout) captured in a CUDA graph.
    At the reference's batch (10 frames x 30 patches) a step is ~25-60 short kernels plus as many allocator calls:
    launch-bound. The graph replays them with one launch; inputs are copied into the captured buffers. Same kernels,
    same order, same arithmetic as the eager step (tests/test_gpu_training.py compares them bit for bit).

        g = trainer.graphed_step(x, top, masks)        # captures; performs NO update
        loss = g(x, masks)                             # one SGD step; `loss` is overwritten by the next call"""

    def __init__(self, trainer, x, top, keep_masks, add_masks=None, mask_rows=None, da_mode=False):
        self.trainer = trainer
        self.x = x.to(torch.float32).contiguous().clone()
        self.keep = None if keep_masks is None else [None if m is None else m.clone() for m in keep_masks]
        self.add = None if add_masks is None else [None if m is None else m.clone() for m in add_masks]
        args = dict(mask_rows=mask_rows, da_mode=da_mode)
        # warm-up on a side stream without touching the weights: first-launch attributes, allocator sizing
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                trainer.step(self.x, top, self.keep, self.add, apply_update=False, **args)
        torch.cuda.current_stream().wait_stream(side)
        step0 = trainer.global_step
        self.graph = torch.cuda.CUDAGraph()
        # No garbage collection while the stream is capturing: an earlier trainer's cached graphs are cyclic garbage
        # (trainer -> cache -> GraphedStep -> trainer), and a collector run inside the capture would destroy their
        # CUDA graphs there - "operation not permitted when stream is capturing" - and invalidate this capture.
        gc.collect()
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = trainer.step(self.x, top, self.keep, self.add, apply_update=True, **args)
        finally:
            if gc_was_enabled:
                gc.enable()
        trainer.global_step = step0          # capturing executes nothing

    @staticmethod
    def _copy_masks(dst, src):
        if dst is None:
            return
        for d, m in zip(dst, src):
            if d is not None:
                d.copy_(m)

    def __call__(self, x, keep_masks=None, add_masks=None):
        self.x.copy_(x.reshape(self.x.shape))
        if keep_masks is not None:
            self._copy_masks(self.keep, keep_masks)
        if add_masks is not None:
            self._copy_masks(self.add, add_masks)
        self.graph.replay()
        self.trainer.global_step += 1
        return self.loss
