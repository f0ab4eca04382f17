"""Python faces of the C-ABI entry points (include/dlc.h). Inputs/outputs are CUDA torch tensors used purely as
device buffers; every function launches hand-written sm_100a kernels from libdlc.so on torch's current stream.
There is no fallback path: without the library or a B200 these raise."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._cuda import Workspace, ptr, stream_ptr

PRECISIONS = {"fp16": _lib.PREC_FP16, "fp16x2": _lib.PREC_FP16X2, "bf16": _lib.PREC_BF16, "auto": _lib.PREC_AUTO,
              "fp16r": _lib.PREC_FP16_REFINED, "fp16x2a16": _lib.PREC_FP16X2_A16}
PRECISION_NAMES = {v: k for k, v in PRECISIONS.items()}
METRICS = {"cos": _lib.METRIC_COS, "cosine": _lib.METRIC_COS, "dot": _lib.METRIC_DOT, "l2": _lib.METRIC_L2}
_DT = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}

_ws = Workspace()


def precision_code(p):
    if isinstance(p, str):
        return PRECISIONS[p]
    return int(p)


def _check_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise ValueError("expected contiguous CUDA tensors")


# ------------------------------------------------------------------------------------------------ planes
def split_planes(src, group_in=1, group_out=1, need_lo=True):
    """[rows, cols] f32/f64 -> (hi, lo) fp16 planes [rows_out, ld]."""
    _check_cuda(src)
    rows, cols = src.shape
    ld = _lib.plane_ld(cols)
    rows_out = -(-rows // group_in) * group_out
    hi = torch.empty((rows_out, ld), dtype=torch.float16, device=src.device)
    lo = torch.empty_like(hi) if need_lo else None
    _lib.call("dlc_split_planes", ptr(src), _DT[src.dtype], rows, cols, src.stride(0), group_in, group_out, ptr(hi),
              ptr(lo), ld, stream_ptr())
    return hi, lo


def pack_weight_planes(w, n_pad=None, need_lo=True):
    """W [k, n] f32/f64 -> transposed planes [n_pad, ld(k)]."""
    _check_cuda(w)
    k, n = w.shape
    n_pad = n_pad or (-(-n // 32) * 32)
    ld = _lib.plane_ld(k)
    hi = torch.empty((n_pad, ld), dtype=torch.float16, device=w.device)
    lo = torch.empty_like(hi) if need_lo else None
    _lib.call("dlc_pack_weight_planes", ptr(w), _DT[w.dtype], k, n, n_pad, ptr(hi), ptr(lo), ld, stream_ptr())
    return hi, lo


def gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, bias=None, act="none", precision="fp16x2", want_f32=True,
                want_planes=False):
    """act(A B^T + bias). Returns (out_f32 [m, n] or None, (out_hi, out_lo) or None)."""
    _check_cuda(a_hi, a_lo, b_hi, b_lo, bias)
    n_pad, ld = b_hi.shape
    assert a_hi.shape[1] == ld and a_hi.shape[0] >= m
    prec = precision_code(precision)
    out = torch.empty((m, n), dtype=torch.float32, device=a_hi.device) if want_f32 else None
    o_hi = o_lo = None
    pld = 0
    if want_planes:
        pld = max(_lib.plane_ld(n), n_pad)
        o_hi = torch.zeros((m, pld), dtype=torch.float16, device=a_hi.device)
        o_lo = torch.zeros_like(o_hi) if prec == _lib.PREC_FP16X2 else None
    acts = {"none": _lib.ACT_NONE, "sigmoid": _lib.ACT_SIGMOID, "relu": _lib.ACT_RELU}
    _lib.call("dlc_gemm_planes", ptr(a_hi), ptr(a_lo), ptr(b_hi), ptr(b_lo), m, n, n_pad, ld, ptr(bias), acts[act],
              prec, ptr(out), n, ptr(o_hi), ptr(o_lo), pld, stream_ptr())
    return out, ((o_hi, o_lo) if want_planes else None)


def matmul(a, b, bias=None, act="none", precision="fp16x2"):
    """a [m, k] @ b [k, n] (+ bias, activation) for f32/f64 CUDA tensors -> float32 [m, n] (tensor cores)."""
    _check_cuda(a, b)
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    split = precision_code(precision) == _lib.PREC_FP16X2
    a_hi, a_lo = split_planes(a, need_lo=split)
    b_hi, b_lo = pack_weight_planes(b, need_lo=split)
    bias_f = None
    if bias is not None:
        bias_f = torch.zeros(b_hi.shape[0], dtype=torch.float32, device=a.device)
        bias_f[:n] = bias.to(torch.float32)
    out, _ = gemm_planes(a_hi, a_lo, b_hi, b_lo, m, n, bias_f, act, precision)
    return out


# ------------------------------------------------------------------------------------------------ patches
SURF_CANDIDATE_CAP = 16384
_ws_surf = Workspace()


def surf_detect(img, top_n=30, hessian_threshold=100.0, n_octaves=4, n_layers=3, chunk=64):
    """img uint8 [B,H,W] (CUDA) -> (xy float32 [B,top_n,2], info float32 [B,top_n,2] = (size, response),
    found int32 [B]): the top_n fast-Hessian keypoints of every frame, best response first (see dlc_surf_detect).
    Frames are processed `chunk` at a time to bound the workspace (~1.7 MB per 640x480 frame)."""
    _check_cuda(img)
    assert img.dtype == torch.uint8 and img.dim() == 3
    B, H, W = img.shape
    xy = torch.empty((B, top_n, 2), dtype=torch.float32, device=img.device)
    info = torch.empty((B, top_n, 2), dtype=torch.float32, device=img.device)
    found = torch.empty((B,), dtype=torch.int32, device=img.device)
    for b0 in range(0, B, chunk):
        n = min(chunk, B - b0)
        ws, ws_bytes = _ws_surf.get(_lib.call("dlc_surf_workspace_bytes", n, H, W, n_octaves, n_layers))
        _lib.call("dlc_surf_detect", ptr(img[b0:b0 + n]), n, H, W, float(hessian_threshold), n_octaves, n_layers, top_n,
                  ptr(xy[b0:b0 + n]), ptr(info[b0:b0 + n]), ptr(found[b0:b0 + n]), ws, ws_bytes, stream_ptr())
    return xy, info, found


def patch_gather(img, xy, patch=41, swap_xy_quirk=True, need_lo=True):
    """img uint8 [B,H,W], xy float32 [B,P,2] -> (hi, lo) planes [B*P, ld(patch^2)]."""
    _check_cuda(img, xy)
    B, H, W = img.shape
    P = xy.shape[1]
    ld = _lib.plane_ld(patch * patch)
    hi = torch.empty((B * P, ld), dtype=torch.float16, device=img.device)
    lo = torch.empty_like(hi) if need_lo else None
    _lib.call("dlc_patch_gather", ptr(img), B, H, W, ptr(xy), P, patch, int(bool(swap_xy_quirk)), ptr(hi), ptr(lo), ld,
              stream_ptr())
    return hi, lo


def patch_gather_u8(img, xy, patch=41, swap_xy_quirk=True):
    """img uint8 [B,H,W], xy float32 [B,P,2] -> ONE fp16 plane [B*P, ld(patch^2)] of raw pixel values 0..255 (exact
    in fp16), the input of an SdaEncoder(input_u8=True)."""
    _check_cuda(img, xy)
    B, H, W = img.shape
    P = xy.shape[1]
    ld = _lib.plane_ld(patch * patch)
    out = torch.empty((B * P, ld), dtype=torch.float16, device=img.device)
    _lib.call("dlc_patch_gather_u8", ptr(img), B, H, W, ptr(xy), P, patch, int(bool(swap_xy_quirk)), ptr(out), ld,
              stream_ptr())
    return out


def patch_gather_f64(img, xy, patch=41, swap_xy_quirk=True):
    _check_cuda(img, xy)
    B, H, W = img.shape
    P = xy.shape[1]
    out = torch.empty((B * P, patch * patch), dtype=torch.float64, device=img.device)
    _lib.call("dlc_patch_gather_f64", ptr(img), B, H, W, ptr(xy), P, patch, int(bool(swap_xy_quirk)), ptr(out),
              stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ SDA encoder
class SdaEncoder:
    """Owner of a dlc_sda handle (packed weights live on the device)."""

    def __init__(self, dims, precision="fp16x2", input_u8=False):
        """input_u8: the planes given to encode_planes hold raw pixel values (patch_gather_u8); the /255 of the
        reference's parser is folded into layer 0, which then needs two tensor-core products instead of three."""
        self.dims = [int(d) for d in dims]
        self.precision = precision
        self.input_u8 = bool(input_u8)
        self._h = C.c_void_p()
        arr = (C.c_int * len(self.dims))(*self.dims)
        _lib.call("dlc_sda_create", C.byref(self._h), len(self.dims) - 1, arr, precision_code(precision))
        if self.input_u8:
            _lib.call("dlc_sda_set_input_u8", self._h, 1)
        self._ws = Workspace()

    def set_layer(self, l, w, b):
        w = np.ascontiguousarray(w, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        if w.shape != (self.dims[l], self.dims[l + 1]) or b.shape != (self.dims[l + 1],):
            raise ValueError("layer %d expects W %s and b %s" % (l, (self.dims[l], self.dims[l + 1]), (self.dims[l + 1],)))
        _lib.call("dlc_sda_set_layer", self._h, l, w.ctypes.data, b.ctypes.data)

    def encode_planes(self, x_hi, x_lo, rows, out=None):
        _check_cuda(x_hi, x_lo)
        if out is None:
            out = torch.empty((rows, self.dims[-1]), dtype=torch.float32, device=x_hi.device)
        need = _lib.call("dlc_sda_workspace_bytes", self._h, rows)
        ws, ws_bytes = self._ws.get(need)
        _lib.call("dlc_sda_encode", self._h, ptr(x_hi), ptr(x_lo), rows, ptr(out), ws, ws_bytes, stream_ptr())
        return out

    def probe_planes(self, x_hi, x_lo, rows):
        """precision="auto": choose the arithmetic now from a sample of this input (synchronises; see dlc_sda_probe).
        encode_planes does it implicitly on its first call after set_layer."""
        ws, ws_bytes = self._ws.get(_lib.call("dlc_sda_workspace_bytes", self._h, rows))
        _lib.call("dlc_sda_probe", self._h, ptr(x_hi), ptr(x_lo), rows, ws, ws_bytes, stream_ptr())
        return self.chosen_precision()

    def chosen_precision(self):
        """Name of the arithmetic encode uses ("fp16" = one tensor product, "fp16x2a16" = two, "fp16x2" = three);
        "unprobed" for an auto encoder before its first encode."""
        return PRECISION_NAMES.get(int(_lib.call("dlc_sda_chosen_precision", self._h)), "unprobed")

    def probe_stats(self):
        """(one-product, two-product) max relative descriptor error against three products on the probe sample."""
        out = (C.c_double * 2)()
        _lib.call("dlc_sda_probe_stats", self._h, C.cast(out, C.c_void_p))
        return float(out[0]), float(out[1])

    def needs_lo_input(self):
        return precision_code(self.precision) in (_lib.PREC_FP16X2, _lib.PREC_AUTO) and not self.input_u8

    def encode(self, x):
        """x [rows, in] f32/f64 CUDA -> float32 [rows, out]."""
        split = precision_code(self.precision) in (_lib.PREC_FP16X2, _lib.PREC_AUTO)
        hi, lo = split_planes(x, need_lo=split)
        return self.encode_planes(hi, lo, x.shape[0])

    def close(self):
        if self._h:
            _lib.call("dlc_sda_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ SDAV similarity
def sdav_weights(desc, mu=0.5, sigma=0.2):
    """desc float32 [N,P,D] -> float64 [D] distinctive weights."""
    _check_cuda(desc)
    N, P, D = desc.shape
    w = torch.empty(D, dtype=torch.float64, device=desc.device)
    ws, ws_bytes = _ws.get(_lib.call("dlc_sdav_similarity_workspace_bytes", N, P, D))
    _lib.call("dlc_sdav_weights", ptr(desc), N, P, D, float(mu), float(sigma), ptr(w), ws, ws_bytes, stream_ptr())
    return w


def sdav_similarity(desc, mu=0.5, sigma=0.2, a=10.0, b=-10.0, weights=None, precision="auto",
                    full_asymmetric=False, out=None):
    """desc float32 [N,P,D] -> float32 [N,N] score matrix (reference order: i<j mirrored, diagonal -1).
    precision: "auto" (default: one fp16 product + exact second pass over the ambiguous rows when a device-side probe
    expects few of them, else three products), "fp16r" (always the former), "fp16x2" (always three products), "fp16"."""
    _check_cuda(desc, weights)
    N, P, D = desc.shape
    if out is None:
        out = torch.empty((N, N), dtype=torch.float32, device=desc.device)
    ws, ws_bytes = _ws.get(_lib.call("dlc_sdav_similarity_workspace_bytes", N, P, D))
    _lib.call("dlc_sdav_similarity", ptr(desc), N, P, D, float(mu), float(sigma), float(a), float(b), ptr(weights),
              precision_code(precision), int(bool(full_asymmetric)), ptr(out), ws, ws_bytes, stream_ptr())
    return out


def sdav_similarity_part(desc, part, n_parts, mu=0.5, sigma=0.2, a=10.0, b=-10.0, weights=None, precision="auto",
                         out=None):
    """The entries of the [N,N] score matrix owned by `part` of `n_parts` (other entries zero): sum over parts = the
    full matrix of sdav_similarity."""
    _check_cuda(desc, weights)
    N, P, D = desc.shape
    if out is None:
        out = torch.empty((N, N), dtype=torch.float32, device=desc.device)
    ws, ws_bytes = _ws.get(_lib.call("dlc_sdav_similarity_workspace_bytes", N, P, D))
    _lib.call("dlc_sdav_similarity_part", ptr(desc), N, P, D, float(mu), float(sigma), float(a), float(b), ptr(weights),
              precision_code(precision), 0, int(part), int(n_parts), ptr(out), ws, ws_bytes, stream_ptr())
    return out


def sdav_similarity_stats(N, P, D):
    """Diagnostics of the last auto / fp16r similarity call (synchronises): dict with the probe's decision."""
    import ctypes
    out = (ctypes.c_double * 6)()
    ws, _ = _ws.get(_lib.call("dlc_sdav_similarity_workspace_bytes", N, P, D))
    _lib.call("dlc_sdav_similarity_stats", N, P, D, ws, ctypes.cast(out, ctypes.c_void_p), stream_ptr())
    keys = ("use_refine", "margin", "sigma", "flagged_frac_estimate", "flagged_rows", "refined_candidates")
    return dict(zip(keys, list(out)))


# ---- staged form for a sequence split over several GPUs (see dlc_sdav_stage_* in include/dlc.h)
plane_ld = _lib.plane_ld
_ws_stage = Workspace()
_ws_colsum = Workspace()    # its own buffer: in a pipelined stream the column sums of step i+1 run next to step i's Gram


def sdav_stage_stats_bytes(frames_per_part):
    return int(_lib.call("dlc_sdav_stage_stats_bytes", int(frames_per_part)))


def _stage_ws(N, P, D, n_parts):
    return _ws_stage.get(max(_lib.call("dlc_sdav_stage_workspace_bytes", N, P, D, n_parts), 129 * 2 * D * 8))


def sdav_stage_colsum(desc_local, out):
    """desc_local float32 [rows, D] (this rank's descriptor rows) -> out float64 [2 D]: column sums, then column sums
    of squares."""
    _check_cuda(desc_local, out)
    rows, D = desc_local.shape
    ws, ws_bytes = _ws_colsum.get(129 * 2 * D * 8)
    _lib.call("dlc_sdav_stage_colsum", ptr(desc_local), rows, D, ptr(out), ws, ws_bytes, stream_ptr())


def sdav_stage_weights(colsums, rows_total, w, centre, mu=0.5, sigma=0.2):
    """colsums float64 [n_parts, 2 D] (all-gathered) -> w float64 [D], centre float32 [D] (the planes' centring
    vector: the dataset mean or zero)."""
    _check_cuda(colsums, w, centre)
    _lib.call("dlc_sdav_stage_weights", ptr(colsums), colsums.shape[0], int(rows_total), colsums.shape[1] // 2, float(mu),
              float(sigma), ptr(w), ptr(centre), stream_ptr())


def sdav_stage_prepare(desc_local, n_local, frames_per_part, P, w, mean, precision, plane_local, plane_lo_local,
                       stats_local):
    """This rank's block: desc_local float32 [frames_per_part*P, D] (first n_local frames valid) -> centred fp16
    plane slice [frames_per_part*P, ld] (+ residual plane for fp16x2) and the stats block (uint8)."""
    _check_cuda(desc_local, w, mean, plane_local, plane_lo_local, stats_local)
    D = desc_local.shape[1]
    _lib.call("dlc_sdav_stage_prepare", ptr(desc_local), int(n_local), int(frames_per_part), int(P), D, ptr(w), ptr(mean),
              precision_code(precision), ptr(plane_local), ptr(plane_lo_local), ptr(stats_local), stream_ptr())


def sdav_stage_gram(plane_all, plane_lo_all, stats_all, n_parts, frames_per_part, N, P, D, precision, part, S,
                    a=10.0, b=-10.0, full_asymmetric=False):
    _check_cuda(plane_all, plane_lo_all, stats_all, S)
    ws, ws_bytes = _stage_ws(N, P, D, n_parts)
    _lib.call("dlc_sdav_stage_gram", ptr(plane_all), ptr(plane_lo_all), ptr(stats_all), int(n_parts),
              int(frames_per_part), int(N), int(P), int(D), float(a), float(b), precision_code(precision),
              int(bool(full_asymmetric)), int(part), ptr(S), ws, ws_bytes, stream_ptr())


def sdav_stage_fix(plane_all, desc_all, N, P, D, precision, part, n_parts, S, a=10.0, b=-10.0, full_asymmetric=False):
    _check_cuda(plane_all, desc_all, S)
    ws, ws_bytes = _stage_ws(N, P, D, n_parts)
    _lib.call("dlc_sdav_stage_fix", ptr(plane_all), ptr(desc_all), int(N), int(P), int(D), float(a), float(b),
              precision_code(precision), int(bool(full_asymmetric)), int(part), int(n_parts), ptr(S), ws, ws_bytes,
              stream_ptr())


def topk_rows(scores, k, largest=True, exclude_band=-1, cand_idx=None):
    """Per-row top-k of a dense [rows, cols] float32 matrix -> (scores [rows,k] f32, idx [rows,k] i64)."""
    _check_cuda(scores, cand_idx)
    rows, cols = scores.shape
    o_s = torch.empty((rows, k), dtype=torch.float32, device=scores.device)
    o_i = torch.empty((rows, k), dtype=torch.int64, device=scores.device)
    _lib.call("dlc_topk_rows", ptr(scores), ptr(cand_idx), rows, cols, scores.stride(0), k, int(bool(largest)),
              int(exclude_band), ptr(o_s), ptr(o_i), stream_ptr())
    return o_s, o_i


def mean_pool_rows(x, group_rows):
    """[groups*group_rows, cols] float32 -> [groups, cols] (frame-level descriptor = mean of its patch descriptors)."""
    _check_cuda(x)
    rows, cols = x.shape
    groups = rows // group_rows
    out = torch.empty((groups, cols), dtype=torch.float32, device=x.device)
    _lib.call("dlc_mean_pool_rows", ptr(x), groups, group_rows, cols, ptr(out), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ Hamming
def hamming_matrix(desc, signed_bin_quirk=True):
    """desc int8 [N,M] -> int32 [N,N]."""
    _check_cuda(desc)
    assert desc.dtype == torch.int8
    N, M = desc.shape
    out = torch.empty((N, N), dtype=torch.int32, device=desc.device)
    ws, ws_bytes = _ws.get(_lib.call("dlc_hamming_workspace_bytes", N, M))
    _lib.call("dlc_hamming_matrix", ptr(desc), N, M, int(bool(signed_bin_quirk)), ptr(out), ws, ws_bytes,
              stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ matrix -> image
IMG_SIMILARITY, IMG_DISTANCE = 0, 1


def matrix_image(m, mode, truncate_int=False):
    """Score matrix (float32) or Hamming matrix (int32) [rows, cols] on the device -> uint8 image, the tail of
    create_similarity_matrix.py:41-48 / create_distance_matrix.py:40-41 (see dlc_matrix_image)."""
    _check_cuda(m)
    if m.dtype == torch.float32:
        dtype = _lib.F32
    elif m.dtype == torch.int32:
        dtype = _lib.I32
    else:
        raise ValueError("expected a float32 or int32 matrix")
    rows, cols = m.shape
    out = torch.empty((rows, cols), dtype=torch.uint8, device=m.device)
    ws, ws_bytes = _ws.get(_lib.call("dlc_matrix_image_workspace_bytes"))
    _lib.call("dlc_matrix_image", ptr(m), dtype, rows, cols, int(mode), int(bool(truncate_int)), ptr(out), ws,
              ws_bytes, stream_ptr())
    return out


# ------------------------------------------------------------------------------------------------ fused conv head
class CnnVtlHead:
    """Owner of a dlc_cnnvtl handle: packed conv filters and the kept-column tables live on the device."""

    def __init__(self, height, width, precision="fp16x2"):
        self.precision = precision
        self._h = C.c_void_p()
        _lib.call("dlc_cnnvtl_create", C.byref(self._h), int(height), int(width), precision_code(precision))
        self.descriptor_len = int(_lib.call("dlc_cnnvtl_descriptor_len", self._h))
        self.n_keep = 0
        self._ws = Workspace()

    def set_conv(self, layer, w, b):
        w = np.ascontiguousarray(w, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        _lib.call("dlc_cnnvtl_set_conv", self._h, layer, w.ctypes.data, b.ctypes.data)

    def set_keep_cols(self, keep_cols):
        keep = np.ascontiguousarray(keep_cols, dtype=np.int64)
        _lib.call("dlc_cnnvtl_set_keep_cols", self._h, keep.ctypes.data, int(keep.size))
        self.n_keep = int(keep.size)

    def forward(self, x, layer_outputs=None, quantise=True):
        """x CUDA [n, H, W, 3] uint8 / float32 / float64 -> int8 [n, M]. layer_outputs: optional list of 5 float32
        CUDA tensors (or None entries) receiving the conv outputs."""
        _check_cuda(x)
        n = x.shape[0]
        out = torch.empty((n, self.n_keep), dtype=torch.int8, device=x.device) if quantise else None
        segs = None
        if layer_outputs is not None:
            _check_cuda(*layer_outputs)
            segs = (C.c_void_p * 5)(*[ptr(t) for t in layer_outputs])
        dt = {torch.uint8: _lib.U8, torch.float32: _lib.F32, torch.float64: _lib.F64}[x.dtype]
        ws, ws_bytes = self._ws.get(_lib.call("dlc_cnnvtl_workspace_bytes", self._h, n))
        _lib.call("dlc_cnnvtl_forward", self._h, ptr(x), dt, n, ptr(out), segs, ws, ws_bytes, stream_ptr())
        return out

    def close(self):
        if self._h:
            _lib.call("dlc_cnnvtl_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ conv blocks
def im2col_planes(x_hi, x_lo, N, H, W, C_, KH, KW, stride, pad_t, pad_l, OH, OW):
    _check_cuda(x_hi, x_lo)
    ld = _lib.plane_ld(KH * KW * C_)
    o_hi = torch.empty((N * OH * OW, ld), dtype=torch.float16, device=x_hi.device)
    o_lo = torch.empty_like(o_hi) if x_lo is not None else None
    _lib.call("dlc_im2col_planes", ptr(x_hi), ptr(x_lo), N, H, W, C_, x_hi.shape[1], KH, KW, stride, pad_t, pad_l, OH,
              OW, ptr(o_hi), ptr(o_lo), ld, stream_ptr())
    return o_hi, o_lo


def maxpool_planes(x, N, H, W, C_, window, stride, need_lo=True):
    _check_cuda(x)
    OH, OW = (H - window) // stride + 1, (W - window) // stride + 1
    ld = _lib.plane_ld(C_)
    o_hi = torch.empty((N * OH * OW, ld), dtype=torch.float16, device=x.device)
    o_lo = torch.empty_like(o_hi) if need_lo else None
    _lib.call("dlc_maxpool_planes", ptr(x), N, H, W, C_, window, stride, OH, OW, ptr(o_hi), ptr(o_lo), ld,
              stream_ptr())
    return o_hi, o_lo, OH, OW


def cnnvtl_quantise(segments, N, keep_cols):
    """segments: list of float32 CUDA tensors [N, size_l]; keep_cols int64 CUDA [M] -> int8 [N, M]."""
    _check_cuda(keep_cols, *segments)
    n_seg = len(segments)
    ptrs = (C.c_void_p * n_seg)(*[s.data_ptr() for s in segments])
    sizes = (C.c_int64 * n_seg)(*[s.numel() // N for s in segments])
    M = keep_cols.numel()
    mm = torch.empty((N, 2), dtype=torch.float32, device=keep_cols.device)
    out = torch.empty((N, M), dtype=torch.int8, device=keep_cols.device)
    _lib.call("dlc_cnnvtl_quantise", ptrs, sizes, n_seg, N, ptr(keep_cols), M, ptr(mm), ptr(out), stream_ptr())
    return out
