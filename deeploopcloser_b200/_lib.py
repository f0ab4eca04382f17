"""ctypes binding of libdlc.so (the C ABI declared in include/dlc.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised. The library is
built in-tree by ``__graft_entry__.build()`` / ``make -C deeploopcloser_b200/csrc``."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DLC_LIB_PATH") or os.path.join(_HERE, "libdlc.so")  # env override: developer A/B builds

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
PREC_FP16, PREC_FP16X2, PREC_BF16, PREC_AUTO, PREC_FP16_REFINED, PREC_FP16X2_A16 = 0, 1, 2, 3, 4, 5
F32, F64, F16, BF16, U8, I32 = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_SIGMOID, ACT_RELU = 0, 1, 2
METRIC_COS, METRIC_DOT, METRIC_L2 = 0, 1, 2


class DlcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libdlc error %d: %s" % (code, message))
        self.code = code


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_sz = C.c_size_t
_d = C.c_double
_f = C.c_float

# name -> (restype, argtypes); every int-returning function is error-checked by `call`.
PROTOTYPES = {
    "dlc_last_error": (C.c_char_p, []),
    "dlc_version": (_i, []),
    "dlc_device_check": (_i, []),
    "dlc_sm_count": (_i, []),
    "dlc_debug_set": (_i, [_i, _i]),
    "dlc_set_sm_reserve": (_i, [_i]),
    "dlc_sdav_debug_gram_only": (_i, [_i]),
    "dlc_plane_ld": (_i, [_i]),
    "dlc_split_planes": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "dlc_pack_weight_planes": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "dlc_gemm_planes": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _p, _p, _i, _p]),
    "dlc_patch_gather": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p, _p, _i, _p]),
    "dlc_patch_gather_u8": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p, _i, _p]),
    "dlc_surf_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dlc_surf_detect": (_i, [_p, _i, _i, _i, _f, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "dlc_patch_gather_f64": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p, _p]),
    "dlc_sda_create": (_i, [C.POINTER(_p), _i, C.POINTER(_i), _i]),
    "dlc_sda_destroy": (_i, [_p]),
    "dlc_sda_set_input_u8": (_i, [_p, _i]),
    "dlc_sda_set_layer": (_i, [_p, _i, _p, _p]),
    "dlc_sda_workspace_bytes": (_sz, [_p, _i]),
    "dlc_sda_encode": (_i, [_p, _p, _p, _i, _p, _p, _sz, _p]),
    "dlc_sda_probe": (_i, [_p, _p, _p, _i, _p, _sz, _p]),
    "dlc_sda_chosen_precision": (_i, [_p]),
    "dlc_sda_probe_stats": (_i, [_p, _p]),
    "dlc_sdav_similarity_workspace_bytes": (_sz, [_i, _i, _i]),
    "dlc_sdav_similarity": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _i, _i, _p, _p, _sz, _p]),
    "dlc_sdav_similarity_part": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "dlc_sdav_similarity_stats": (_i, [_i, _i, _i, _p, _p, _p]),
    "dlc_sdav_weights": (_i, [_p, _i, _i, _i, _d, _d, _p, _p, _sz, _p]),
    "dlc_sdav_stage_stats_bytes": (_sz, [_i]),
    "dlc_sdav_stage_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dlc_sdav_stage_colsum": (_i, [_p, _i64, _i, _p, _p, _sz, _p]),
    "dlc_sdav_stage_weights": (_i, [_p, _i, _i64, _i, _d, _d, _p, _p, _p]),
    "dlc_sdav_stage_prepare": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p]),
    "dlc_sdav_stage_gram": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _d, _d, _i, _i, _i, _p, _p, _sz, _p]),
    "dlc_sdav_stage_fix": (_i, [_p, _p, _i, _i, _i, _d, _d, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "dlc_topk_rows": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "dlc_mean_pool_rows": (_i, [_p, _i, _i, _i, _p, _p]),
    "dlc_db_create": (_i, [C.POINTER(_p), _i, _i64, _i, _i]),
    "dlc_db_destroy": (_i, [_p]),
    "dlc_db_size": (_i64, [_p]),
    "dlc_db_clear": (_i, [_p]),
    "dlc_db_append": (_i, [_p, _p, _i, _i64, _p]),
    "dlc_match_workspace_bytes": (_sz, [_p, _i, _i]),
    "dlc_match_topk": (_i, [_p, _p, _i, _i, _i64, _p, _p, _p, _sz, _p]),
    "dlc_match_threshold": (_i, [_p, _p, _i, _f, _i, _i64, _p, _p, _p, _p, _sz, _p]),
    "dlc_comm_unique_id": (_i, [_p]),
    "dlc_comm_create": (_i, [C.POINTER(_p), _p, _i, _i]),
    "dlc_comm_destroy": (_i, [_p]),
    "dlc_match_sharded_workspace_bytes": (_sz, [_p, _i, _i, _i]),
    "dlc_match_topk_sharded": (_i, [_p, _p, _p, _i, _i, _i64, _p, _p, _p, _sz, _p]),
    "dlc_hamming_workspace_bytes": (_sz, [_i, _i]),
    "dlc_hamming_matrix": (_i, [_p, _i, _i, _i, _p, _p, _sz, _p]),
    "dlc_matrix_image_workspace_bytes": (_sz, []),
    "dlc_matrix_image": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "dlc_cnnvtl_create": (_i, [C.POINTER(_p), _i, _i, _i]),
    "dlc_cnnvtl_destroy": (_i, [_p]),
    "dlc_cnnvtl_set_conv": (_i, [_p, _i, _p, _p]),
    "dlc_cnnvtl_descriptor_len": (_i64, [_p]),
    "dlc_cnnvtl_set_keep_cols": (_i, [_p, _p, _i]),
    "dlc_cnnvtl_workspace_bytes": (_sz, [_p, _i]),
    "dlc_cnnvtl_forward": (_i, [_p, _p, _i, _i, _p, C.POINTER(_p), _p, _sz, _p]),
    "dlc_train_corrupt": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _i, _p]),
    "dlc_train_xent_grad": (_i, [_p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _i, _p]),
    "dlc_train_hidden_grad": (_i, [_p, _p, _p, _i, _i, _i, _f, _d, _d, _p, _p, _p, _p, _i, _p, _p]),
    "dlc_train_colsum": (_i, [_p, _i, _i, _p, _p]),
    "dlc_train_transpose_planes": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _p]),
    "dlc_train_sgd": (_i, [_p, _p, _i, _i64, _d, _p]),
    "dlc_train_mask_grad": (_i, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "dlc_im2col_planes": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "dlc_maxpool_planes": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "dlc_cnnvtl_quantise": (_i, [C.POINTER(_p), C.POINTER(_i64), _i, _i, _p, _i, _p, _p, _p]),
}
_CHECKED = {n for n, (r, _) in PROTOTYPES.items() if r is _i and n not in ("dlc_version", "dlc_sm_count", "dlc_plane_ld", "dlc_sda_chosen_precision")}

_lib = None


def load():
    """Load libdlc.so (once). Raises if it has not been built - there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libdlc.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C deeploopcloser_b200/csrc` (sm_100a, nvcc). deeploopcloser_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        if os.environ.get("DLC_LIB_PATH") and not hasattr(lib, name):
            continue  # developer A/B build of an older revision
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().dlc_last_error().decode("utf-8", "replace")


def call(name, *args):
    """Call an exported function; int-returning entry points raise DlcError on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _CHECKED and rc != OK:
        raise DlcError(rc, last_error())
    return rc


def plane_ld(cols):
    return (int(cols) + 63) // 64 * 64
