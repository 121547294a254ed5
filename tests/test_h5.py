"""N3: Keras HDF5 weight files (train.py:23 ModelCheckpoint(MODEL_FILE), util.py:19 load_weights) through
h5lite, the package's own reader / writer of the HDF5 subset Keras uses.  No GPU: the Engine is not involved."""
import os
import struct

import numpy as np
import pytest

from music_generator_b200 import h5lite
from music_generator_b200 import keras_like as K
from music_generator_b200.config import ModelConfig, param_shapes

# a file written by the HDF5 library itself (MATLAB 7.4 -v7.3 = HDF5 1.6 behind a 512-byte user block) that ships
# with scipy's test-suite: the one libhdf5-made file in this image
SCIPY_MAT = os.path.join(os.path.dirname(__import__("scipy").__file__), "io", "matlab", "tests", "data",
                         "testhdf5_7.4_GLNX86.mat")


def _state(cfg=ModelConfig(), seed=0):
    rs = np.random.RandomState(seed)
    return {k: rs.randn(*shp).astype(np.float32) for k, shp in param_shapes(cfg).items()}


@pytest.mark.skipif(not os.path.exists(SCIPY_MAT), reason="scipy test data not installed")
def test_reader_parses_a_file_written_by_libhdf5():
    """Superblock behind a user block, root symbol table -> B-tree -> symbol node -> local heap, a version-1 object
    header, dataspace / datatype / layout messages and a fixed-string attribute, as the HDF5 library wrote them."""
    with h5lite.File(SCIPY_MAT) as f:
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        assert np.allclose(d.read().ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
        assert d.attrs["MATLAB_class"].tobytes() == b"double"


def test_keras_save_weights_layout_round_trip(tmp_path):
    """save_weights layout: root attrs layer_names / backend / keras_version, a group per layer with weight_names, the
    variables one level down under their TensorFlow names; all 28 tensors come back bit for bit."""
    want = _state()
    path = str(tmp_path / "model.h5")
    h5lite.write(path, K.keras_h5_tree(want))
    with h5lite.File(path) as f:
        names = [n.decode() for n in f.attrs["layer_names"]]
        assert names == [k for k, _, _ in K.KERAS_LAYERS]
        assert f.attrs["backend"].tobytes() == b"tensorflow"
        g = f["time_distributed_4"]
        assert [n.decode() for n in g.attrs["weight_names"]] == [
            "time_distributed_4/kernel:0", "time_distributed_4/recurrent_kernel:0", "time_distributed_4/bias:0"]
        assert g["time_distributed_4/recurrent_kernel:0"].shape == (256, 1024)
        got = K.read_keras_h5(f, param_shapes(ModelConfig()))
    assert set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)


def test_full_model_file_and_shifted_auto_names(tmp_path):
    """ModelCheckpoint without save_weights_only writes `model.save` files: the same tree below /model_weights, beside
    optimizer state.  Auto-names shift when a session has built other layers before (dense_7 instead of dense_1);
    layers are identified by their variable shapes, so such a file still loads.  Layers without weights are present
    as empty groups, as Keras writes them."""
    want = _state(seed=3)
    tree = K.keras_h5_tree(want)
    shifted = {"@attrs": dict(tree["@attrs"])}
    ren = {}
    for layer, _, _ in K.KERAS_LAYERS:
        base, _, num = layer.rpartition("_")
        ren[layer] = f"{base}_{int(num) + 20}" if num.isdigit() else layer
    shifted["@attrs"]["layer_names"] = [ren[k] for k, _, _ in K.KERAS_LAYERS] + ["dropout_1", "input_1"]
    for layer, _, parts in K.KERAS_LAYERS:
        new = ren[layer]
        sub = tree[layer]
        shifted[new] = {"@attrs": {"weight_names": [f"{new}/{w.split('/')[1]}" for w in sub["@attrs"]["weight_names"]]},
                        new: sub[layer]}
    shifted["dropout_1"] = {}
    shifted["input_1"] = {}
    path = str(tmp_path / "full.h5")
    h5lite.write(path, {"model_weights": shifted, "optimizer_weights": {"iterations:0": np.array(7, np.int64)},
                        "@attrs": {"keras_version": b"2.0.8"}})
    with h5lite.File(path) as f:
        assert "layer_names" not in f.attrs and sorted(f.keys()) == ["model_weights", "optimizer_weights"]
        it = f["optimizer_weights/iterations:0"]
        assert it.shape == () and int(it.read()) == 7
        assert f["model_weights/dropout_1"].keys() == []
        got = K.read_keras_h5(f, param_shapes(ModelConfig()))
    assert all(np.array_equal(got[k], want[k]) for k in want)


def test_scaled_model_shapes_are_identified_too(tmp_path):
    cfg = ModelConfig(time_axis_units=512, note_axis_units=256)
    want = _state(cfg, seed=5)
    path = str(tmp_path / "scaled.h5")
    h5lite.write(path, K.keras_h5_tree(want))
    with h5lite.File(path) as f:
        got = K.read_keras_h5(f, param_shapes(cfg))
    assert all(np.array_equal(got[k], want[k]) for k in want)


def test_file_structure_follows_the_format_specification(tmp_path):
    """The writer cannot be checked against libhdf5 here, so the bytes are checked against the specification's
    fixed points: signature and superblock fields, end-of-file address, every structure 8-byte aligned with its
    signature, symbol-table entries in strcmp order, header message sizes multiples of 8."""
    path = str(tmp_path / "w.h5")
    h5lite.write(path, {"b": {"x": np.arange(6, dtype=np.float32).reshape(2, 3)}, "a": np.int32(5), "B": {},
                        "@attrs": {"names": [b"b", b"a"], "n": np.arange(3, dtype=np.int64)}})
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8:13] == b"\0\0\0\0\0" and raw[13:15] == b"\x08\x08"
    leaf_k, internal_k = struct.unpack_from("<HH", raw, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", raw, 24)
    assert (base, free, drv) == (0, h5lite.UNDEF, h5lite.UNDEF) and eof == len(raw) and len(raw) % 8 == 0
    _, root_hdr, cache, _ = struct.unpack_from("<QQII", raw, 56)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert cache == 1 and raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"
    assert root_hdr % 8 == 0 and btree % 8 == 0 and heap % 8 == 0
    # the root header's symbol-table message repeats the scratch-pad addresses
    ver, _, nmsg, refs, hsize = struct.unpack_from("<BBHII", raw, root_hdr)
    assert (ver, refs) == (1, 1) and hsize % 8 == 0 and nmsg == 3
    mtype, msize = struct.unpack_from("<HH", raw, root_hdr + 16)
    assert mtype == 0x0011 and msize == 16 and struct.unpack_from("<QQ", raw, root_hdr + 24) == (btree, heap)
    # B-tree node: one child, keys bracket it; node sized by the superblock's internal K
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", raw, btree + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, h5lite.UNDEF, h5lite.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", raw, btree + 24)
    assert raw[snod:snod + 4] == b"SNOD" and raw[snod + 4] == 1
    dsize, free_head, daddr = struct.unpack_from("<QQQ", raw, heap + 8)
    assert free_head == 1 and dsize % 8 == 0 and raw[daddr:daddr + 8] == b"\0" * 8 and key0 == 0
    n = struct.unpack_from("<H", raw, snod + 6)[0]
    names = []
    for i in range(n):
        off = struct.unpack_from("<Q", raw, snod + 8 + 40 * i)[0]
        names.append(raw[daddr + off:raw.index(b"\0", daddr + off)])
    assert names == [b"B", b"a", b"b"] and raw[daddr + key1:daddr + key1 + 2] == b"b\0"
    assert len(raw) >= snod + 8 + 40 * 2 * leaf_k and internal_k == 16
    with h5lite.File(path) as f:
        assert [x.decode() for x in f.attrs["names"]] == ["b", "a"] and list(f.attrs["n"]) == [0, 1, 2]
        assert int(f["a"].read()) == 5 and f["B"].keys() == []
        assert np.array_equal(f["b/x"].read(), np.arange(6, dtype=np.float32).reshape(2, 3))


def test_not_hdf5_and_new_style_files_are_clear_errors(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"PK\x03\x04" + b"\0" * 600)                 # an .npz under an .h5 name
    with pytest.raises(h5lite.H5Error, match="not an HDF5 file"):
        h5lite.File(str(p))
    p.write_bytes(b"\x89HDF\r\n\x1a\n\x02" + b"\0" * 600)        # superblock version 2 (libver='latest')
    with pytest.raises(h5lite.H5Error, match="superblock version 2"):
        h5lite.File(str(p))


def test_reader_handles_chunked_deflate_shuffle_datasets(tmp_path):
    """Keras never writes chunked weights, other producers do (h5py with compression=...): data layout class 2 through
    the version-1 raw-data B-tree, with the shuffle and deflate filters.  The file is assembled here from the
    specification's structures (the writer only emits contiguous datasets)."""
    import zlib
    w = h5lite._Writer()
    a = (np.arange(7 * 10, dtype=np.float32).reshape(7, 10) * 0.5 - 3).astype("<f4")
    chunk = (4, 8)
    keys = []
    for r0 in range(0, 7, chunk[0]):
        for c0 in range(0, 10, chunk[1]):
            tile = np.zeros(chunk, "<f4")
            blk = a[r0:r0 + chunk[0], c0:c0 + chunk[1]]
            tile[:blk.shape[0], :blk.shape[1]] = blk
            raw = np.frombuffer(tile.tobytes(), np.uint8).reshape(-1, 4).T.tobytes()      # shuffle (filter 2)
            raw = zlib.compress(raw)                                                       # deflate (filter 1)
            keys.append((len(raw), (r0, c0, 0), w.alloc(raw)))
    node = bytearray(b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), h5lite.UNDEF, h5lite.UNDEF))
    for size, offs, addr in keys:
        node += struct.pack("<II", size, 0) + struct.pack("<QQQ", *offs) + struct.pack("<Q", addr)
    node += struct.pack("<II", 0, 0) + struct.pack("<QQQ", 8, 0, 0)                        # the closing key
    btree = w.alloc(bytes(node))
    filt = struct.pack("<BB6x", 1, 2)                                                      # pipeline v1, two filters
    filt += struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 4, 0)                    # shuffle, 1 client value (+pad)
    filt += struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 6, 0)                    # deflate level 6 (+pad)
    msgs = [(0x0001, h5lite._space_msg(a.shape)), (0x0003, h5lite._dtype_msg(a.dtype)),
            (0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)), (0x000B, filt),
            (0x0008, struct.pack("<BBBQIII", 3, 2, 3, btree, chunk[0], chunk[1], 4))]
    dset = w.alloc(h5lite._header(msgs))
    # a root group holding that one dataset: reuse the writer's group builder with a placeholder, then patch the link
    root = w.group({"x": np.zeros(1, np.float32)})
    data = bytearray(w.finish(root))
    snod = data.index(b"SNOD")
    struct.pack_into("<Q", data, snod + 8 + 8, dset)                                        # entry 0: object header address
    p = tmp_path / "chunked.h5"
    p.write_bytes(bytes(data))
    with h5lite.File(str(p)) as f:
        d = f["x"]
        assert d.shape == (7, 10) and d._filters() == [(2, (4,)), (1, (6,))]
        assert np.array_equal(d.read(), a)
