"""CPU: host-side mirror of the reference API (no kernel launches) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------- C ABI
def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "dlc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dlc_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from deeploopcloser_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libdlc.so does not export %s" % n
        assert n in _lib.PROTOTYPES, "no ctypes prototype for %s" % n
    assert lib.dlc_version() >= 100
    assert lib.dlc_plane_ld(1681) == 1728 and lib.dlc_plane_ld(2500) == 2560 and lib.dlc_plane_ld(64) == 64


def test_no_cpu_fallback():
    """Without a GPU the product raises instead of computing something on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deeploopcloser_b200.da import DA
    from deeploopcloser_b200.distance import DistanceCalculator
    with pytest.raises(RuntimeError):
        DA([30, 64], 32).transform(np.zeros((30, 64)))
    with pytest.raises(RuntimeError):
        DistanceCalculator.calculate_distance(np.zeros(8, np.int8), np.ones(8, np.int8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deeploopcloser_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


# ------------------------------------------------------------------------------------------- TensorWrapper
def test_tensorwrapper_reference_test_graph():
    """Graph construction of the reference's only test (test/TensorflowWrapperTest.py:11-16): a 3-D x 2-D matmul is
    flatten -> matmul -> re-batch (TensorflowWrapper.py:57-67)."""
    import src.utils.TensorflowWrapper as tw
    x = tw.constant([[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]])
    w = tw.constant([[2, 2], [2, 2]])
    assert x.dimensions() == 3 and w.dimensions() == 2 and x.to_tf().dtype is np.float64
    y = x.matmul(w.to_tf()).to_tf()
    assert y.op == "reshape" and y.static_rank == 3
    mm = y.inputs[0]
    assert mm.op == "matmul" and mm.inputs[0].op == "reshape" and mm.inputs[0].static_rank == 2
    assert tw.parameter_guard([x, 3]) == [x.to_tf(), 3]
    assert x.batch_size().to_tf().op == "getitem" and w.batch_size() == 1


def test_tensorwrapper_encoder_layer_pattern():
    import src.utils.TensorflowWrapper as tw
    x0 = tw.placeholder(tw.float64, [None, 30, 1681])
    h0 = x0.corrupt(0).matmul(tw.constant(np.zeros((1681, 8)))).add(tw.constant(np.zeros(8))).sigmoid()
    n = h0.to_tf()
    assert n.op == "sigmoid" and n.inputs[0].op == "add"
    assert h0.dimensions() == 3


# ------------------------------------------------------------------------------------------- SDAV / DA / SDA
def test_sdav_surface():
    from src.sdav.network.SDAV import SDAV
    m = SDAV.__new__(SDAV)
    m._define_params()
    assert m.input_shape == [30, 1681] and m.hidden_units == [2500] * 5 and m.default_batch_size == 10
    assert (m.corruption_level, m.sparse_level, m.sparse_penalty, m.consecutive_penalty) == (0.3, 0.05, 1.0, 0.2)
    assert m.learning_rate == 0.1 and m.epochs == 100
    assert m.get_layer_input_shape(0) == [30, 1681] and m.get_layers_input_shapes() == [[30, 2500]] * 5
    for name in ("transform", "transform_dataset", "transform_all", "fit", "fit_dataset", "get_dataset"):
        assert callable(getattr(SDAV, name))


def test_sdav_weight_roundtrip(tmp_path):
    from deeploopcloser_b200.sdav import SDAV
    m = SDAV.__new__(SDAV)
    m._define_params()
    m.hidden_units = [16, 8]
    m.input_shape = [30, 25]
    m._set_train_path(str(tmp_path))
    m._encoder = None
    m.init_weights(3)
    p = m.save_weights()
    w0 = m._weights[0].copy()
    m.init_weights(4)
    assert not np.array_equal(m._weights[0], w0)
    m.load_weights(p)
    assert np.array_equal(m._weights[0], w0) and m._weights[1].shape == (16, 8)
    with pytest.raises(ValueError):
        m.set_weights([w0], [np.zeros(16)])


def test_da_sda_validation():
    from src.sdav.network.DenoisingAutoencoderVariant import DA
    from src.sdav.network.StackedDenoisingAutoencoderVariants import SDA
    with pytest.raises(ValueError):
        DA([30, 64], 32, learning_rate=-0.1)
    with pytest.raises(ValueError):
        DA([30], 32)
    with pytest.raises(ValueError):
        SDA([30, 64], [32])
    s = SDA([30, 64], [32, 16], sparse_penalty=1)
    assert [l.input_shape for l in s._layers] == [[30, 64], [30, 32]] and [l.layer_n for l in s._layers] == [0, 1]
    import torch
    if not torch.cuda.is_available():          # training runs on the B200 only: no CPU fallback, it fails loudly
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            s.fit_dataset([np.zeros((30, 64))] * 10)


# ------------------------------------------------------------------------------------------- input parser
def test_boundaries_match_oracle():
    from oracle import patches as o_patch
    from src.sdav.input.CvInputParser import get_1d_boundaries, get_2d_boundaries
    rng = np.random.default_rng(0)
    pts = rng.integers(-5, 260, (200, 2))
    for axis, L in ((0, 192), (1, 240)):
        lo, hi = get_1d_boundaries([192, 240], pts, 41, axis)
        olo, ohi = o_patch.window_bounds(L, pts[:, axis], 41)
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    assert len(get_2d_boundaries([192, 240], pts, 41)) == 4
    with pytest.raises(ValueError):
        get_1d_boundaries([192, 240], pts, 40, 0)
    with pytest.raises(ValueError):
        get_1d_boundaries([192], pts, 41, 0)


# ------------------------------------------------------------------------------------------- CnnVtl
def test_cnnvtl_geometry_and_mask():
    from deeploopcloser_b200 import cnn_vtl as c
    from oracle import cnnvtl as o_cnn
    net = c.CnnVtl(input_shape=[4, 192, 240, 3], weights="synthetic", seed=4)
    assert net.layer_sizes == o_cnn.layer_sizes((192, 240))
    assert net.keep_cols.size <= 2243 and np.all(np.diff(net.keep_cols) > 0)
    assert c.layer_geometry(192, 240)[0][:4] == (192, 240, 46, 58) and c.layer_geometry(192, 240)[1][4:] == (2, 2)
    # default 224x224 input: conv1 'valid' 11x11/4 -> 54x54 (TF floor rule), pools 3/2 -> 26 -> 12
    assert c.CnnVtl((1, 224, 224, 3), weights="synthetic", seed=0).layer_sizes == [279936, 173056, 55296, 55296, 36864]
    mask = np.zeros(546944, bool)
    mask[[5, 300000, 546943]] = True
    assert c.CnnVtl([1, 192, 240, 3], weights="synthetic", mask=mask).keep_cols.tolist() == [5, 300000, 546943]
    with pytest.raises(ValueError):
        c.CnnVtl([1, 192, 240, 3])  # weights are mandatory (the reference's blob is an LFS pointer)
    assert np.array_equal(c._flat_fill([1, 2, 3], (2, 3)), [[1, 2, 3], [3, 3, 3]])
    w = c.synthetic_weights(3)
    ow = o_cnn.make_weights(3)
    assert all(np.array_equal(w[k][0], ow[k][0]) for k in w)


def test_mathutils_golden(golden_dir):
    from src.utils.MathUtils import MathUtils
    g = np.load(golden_dir + "/misc.npz")
    assert [MathUtils.compressed_size(int(v), 99.59) for v in g["vals"]] == list(g["compressed"])


def test_tf_checkpoint_roundtrip_and_sdav_restore(tmp_path):
    """Weight ingestion from the reference's formats: a TensorFlow V2 checkpoint with the SDAV
    graph's variable names round-trips bit-exactly, crc32c is verified, the `checkpoint` state file resolves the latest
    prefix like tf.train.latest_checkpoint, and the SDAV class restores from the directory."""
    from deeploopcloser_b200 import tf_checkpoint as tc
    assert tc.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 64 * (1 << 14) + 5, dtype=np.uint8).tobytes()
    assert tc._crc32c_np(raw) == tc.crc32c(raw)                          # vectorised CRC == bytewise CRC
    dims = [7, 5, 4]
    names = tc.sdav_variable_names(2)
    assert names[0] == ("Variable", "Variable_1", "Variable_2") and names[1][0] == "Variable_3"
    tensors = {"global_step": np.array(12, dtype=np.int32)}
    for l, (wn, bn, dn) in enumerate(names):
        tensors[wn] = rng.standard_normal((dims[l], dims[l + 1]))
        tensors[bn] = rng.standard_normal(dims[l + 1])
        tensors[dn] = rng.standard_normal(dims[l])
    prefix = tc.save_checkpoint(str(tmp_path / "checkpoint_file-12"), tensors)
    assert tc.latest_checkpoint(str(tmp_path)) == prefix
    assert ("Variable_3", (5, 4), np.dtype("float64")) in tc.list_variables(prefix)
    back = tc.load_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape and np.array_equal(back[k], tensors[k])
    # a flipped byte in the data file is caught by the stored crc32c
    data = tmp_path / "checkpoint_file-12.data-00000-of-00001"
    blob = bytearray(data.read_bytes())
    blob[3] ^= 0x40
    data.write_bytes(bytes(blob))
    with pytest.raises(ValueError, match="crc32c"):
        tc.load_checkpoint(prefix)
    blob[3] ^= 0x40
    data.write_bytes(bytes(blob))
    # the SDAV class restores from the checkpoint directory and writes a checkpoint the same reader accepts
    from deeploopcloser_b200.sdav import SDAV
    net = SDAV.__new__(SDAV)
    net.input_shape, net.hidden_units = [3, 7], [5, 4]
    net._encoder = None
    net.checkpoints_path = str(tmp_path / "out")
    net.load_weights(str(tmp_path))
    assert net.global_step == 12 and np.array_equal(net._weights[1], tensors["Variable_3"])
    assert np.array_equal(net._dec_biases[0], tensors["Variable_2"])
    p2 = net.save_weights(fmt="tf")
    assert os.path.basename(p2) == "checkpoint_file-12"
    again = tc.load_checkpoint(tc.latest_checkpoint(net.checkpoints_path))
    assert all(np.array_equal(again[k], tensors[k]) for k in tensors)


def _crc32c_bitwise(data):
    """CRC-32C written independently of the product (bit by bit, no table)."""
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def _hand_block(entries, restarts):
    """Table block assembled by hand from (shared, unshared key bytes, value) triples and explicit restart offsets -
    independent of the product's writer (table_format.md: varint32 shared | non_shared | value_length, key delta, value;
    uint32 restarts[], uint32 num_restarts; trailer = type byte 0 + masked crc32c of block + type)."""
    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)
    body = b"".join(varint(s) + varint(len(k)) + varint(len(v)) + k + v for s, k, v in entries)
    body += b"".join(r.to_bytes(4, "little") for r in restarts) + len(restarts).to_bytes(4, "little")
    c = _crc32c_bitwise(body + b"\x00")
    masked = (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF
    return body, body + b"\x00" + masked.to_bytes(4, "little"), varint


def test_tf_checkpoint_index_with_shared_prefixes_and_several_blocks(tmp_path):
    """Real tf.train.Saver indices are LevelDB tables with prefix-compressed keys (restart interval 16) and may span
    several data blocks. (1) An index assembled by hand from the table-format specification - shared-prefix key
    deltas, a mid-block restart point, two data blocks, bit-wise CRC - is read correctly; a flipped byte in a block is
    caught by the block trailer. (2) The writer emits that layout and round-trips through the `shared > 0` branch."""
    from deeploopcloser_b200 import tf_checkpoint as tc
    # ---- (1) hand-assembled: keys Variable, Variable_1, Variable_10 | (restart) Variable_11 ; second block: global_step
    v = [b"v0", b"value-1", b"v10", b"v11", b"gs"]
    e0 = (0, b"Variable", v[0])
    e1 = (8, b"_1", v[1])                       # shares "Variable"
    e2 = (10, b"0", v[2])                       # shares "Variable_1"
    off_restart = sum(3 + len(k) + len(val) for _, k, val in (e0, e1, e2))       # 1-byte varints: 3 header bytes each
    e3 = (0, b"Variable_11", v[3])              # restart point: full key
    b1_body, b1, varint = _hand_block([e0, e1, e2, e3], [0, off_restart])
    b0_body, b0, _ = _hand_block([(0, b"global_step", v[4])], [0])               # sorts before "Variable..." ('g' > 'V')
    # table: data block A (Variable*), data block B (global_step) - keys ascending: "Variable..." < "global_step"
    blob = bytearray()
    handles = []
    for body, full, last in ((b1_body, b1, b"Variable_11"), (b0_body, b0, b"global_step")):
        handles.append((last, varint(len(blob)) + varint(len(body))))
        blob += full
    meta_body, meta, _ = _hand_block([], [0])
    meta_off = len(blob)
    blob += meta
    idx_body, idx, _ = _hand_block([(0, k, h) for k, h in handles], [0, 3 + len(handles[0][0]) + len(handles[0][1])])
    idx_off = len(blob)
    blob += idx
    footer = varint(meta_off) + varint(len(meta_body)) + varint(idx_off) + varint(len(idx_body))
    blob += footer + b"\x00" * (40 - len(footer)) + (0xdb4775248b80fb57).to_bytes(8, "little")
    path = tmp_path / "hand.index"
    path.write_bytes(bytes(blob))
    got = tc._read_table(str(path))
    assert got == {b"Variable": v[0], b"Variable_1": v[1], b"Variable_10": v[2], b"Variable_11": v[3],
                   b"global_step": v[4]}
    bad = bytearray(blob)
    bad[5] ^= 0x01                              # inside data block A
    path.write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="crc32c"):
        tc._read_table(str(path))
    # ---- (2) the writer: 40 variables -> prefix sharing (restart interval 16) and, with small blocks, several blocks
    rng = np.random.default_rng(5)
    tensors = {("Variable" if i == 0 else "Variable_%d" % i): rng.standard_normal((3, i + 1)) for i in range(40)}
    tensors["global_step"] = np.array(7, dtype=np.int64)
    prefix = tc.save_checkpoint(str(tmp_path / "ck-7"), tensors, block_size=300)
    raw = (tmp_path / "ck-7.index").read_bytes()
    # parse the index block by hand: more than one data block, and shared > 0 entries inside the first one
    footer = raw[-48:]
    _, p = tc._get_varint(footer, 0)
    _, p = tc._get_varint(footer, p)
    idx_off, p = tc._get_varint(footer, p)
    idx_size, _ = tc._get_varint(footer, p)
    blocks = tc._read_block(raw, idx_off, idx_size)
    assert len(blocks) > 3
    off, p = tc._get_varint(blocks[1][1], 0)
    shared, q = tc._get_varint(raw, off)        # first entry of a block is a restart point ...
    non_shared, q = tc._get_varint(raw, q)
    vlen, q = tc._get_varint(raw, q)
    assert shared == 0
    assert tc._get_varint(raw, q + non_shared + vlen)[0] > 0   # ... the next key shares a prefix with it
    back = tc.load_checkpoint(prefix)
    assert set(back) == set(tensors) and all(np.array_equal(back[k], tensors[k]) for k in tensors)
    assert [n for n, _, _ in tc.list_variables(prefix)] == sorted(tensors)
    # restart interval 1 (no sharing) reads the same
    p1 = tc.save_checkpoint(str(tmp_path / "ck1-7"), tensors, update_state=False, restart_interval=1)
    assert all(np.array_equal(tc.load_checkpoint(p1)[k], tensors[k]) for k in tensors)


def test_bench_reference_arm_is_a_measurement():
    """bench.py --impl reference runs a fixed sub-workload to completion and reports measured and extrapolated figures
    separately (both the hoisted and the literal, mean-per-pair variant of the pair loop)."""
    import bench
    m = bench.cpu_step_measured(n_frames=5, unhoisted=True)
    assert m["frames"] == 5 and m["pairs"] == 10
    assert m["measured_step_s"] == pytest.approx(m["encode_s"] + m["pairs_hoisted_s"])
    assert m["extrapolated_full_step_s"] > m["measured_step_s"]
    assert m["extrapolated_full_step_unhoisted_s"] >= m["extrapolated_full_step_s"]
    text = bench.reference_sample_text(m)
    assert "5 frames encoded" in text and "all 10 frame pairs" in text and "literal reference" in text


def test_frame_blocks_cover_the_sequence():
    """ShardedSequencePipeline.frame_block: contiguous blocks of ceil(N / world) frames, the last ones shorter or empty,
    together exactly the sequence (the staged ABI relies on global frame g living in block g // per at g % per)."""
    from deeploopcloser_b200.pipeline import ShardedSequencePipeline as S
    for n, world in ((1063, 8), (1063, 3), (7, 2), (5, 8), (16, 4), (1, 1)):
        seen = []
        per0 = -(-n // world)
        for r in range(world):
            start, end, per = S.frame_block(n, r, world)
            assert per == per0 and 0 <= end - start <= per
            assert start == min(r * per, n)
            seen += list(range(start, end))
        assert seen == list(range(n))


def test_sm_reserve_is_scoped_to_the_sharded_step(monkeypatch):
    """ShardedSequencePipeline.sm_reserve: the process-wide dlc_set_sm_reserve knob is set for the duration of a step and
    restored to 0 afterwards, also when the step raises; 0 / None never touch it; DLC_SM_RESERVE sets the default of a
    multi-rank pipeline (host logic only - no library call is made here)."""
    from deeploopcloser_b200 import pipeline
    calls = []
    monkeypatch.setattr(pipeline._lib, "call", lambda name, *a: calls.append((name,) + a))
    with pipeline._SmReserve(8):
        assert calls == [("dlc_set_sm_reserve", 8)]
    assert calls == [("dlc_set_sm_reserve", 8), ("dlc_set_sm_reserve", 0)]
    calls.clear()
    with pytest.raises(ValueError):
        with pipeline._SmReserve(4):
            raise ValueError("step failed")
    assert calls[-1] == ("dlc_set_sm_reserve", 0)
    calls.clear()
    for n in (0, None):
        with pipeline._SmReserve(n):
            pass
    assert calls == []
    assert pipeline.SM_RESERVE_DEFAULT == 0      # measured on 8 GPUs: no gain beyond the noise
