"""CPU-side checks: the C-ABI library loads and exports every symbol `include/seggroup_b200.h` declares,
argument validation works without a device, the host-side loaders behave."""
import ctypes
import os

import numpy as np
import pytest
import torch


def test_library_exports_every_declared_symbol():
    from seggroup_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 40
    lib = ctypes.CDLL(_lib.LIB_PATH) if os.path.isfile(_lib.LIB_PATH) else None
    _lib.load()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing
    assert _lib.call("sgb_version") == 100


def test_workspace_queries_and_argument_validation():
    from seggroup_b200 import _lib
    assert _lib.call("sgb_scan_ws_bytes", 10) > 0
    assert _lib.call("sgb_segment_pool_ws_bytes", 100, 64) == 100 * 64 * 8
    assert _lib.call("sgb_edgeconv_ws_bytes", 150000, 1) > _lib.call("sgb_edgeconv_ws_bytes", 150000, 0)
    with pytest.raises(_lib.SgbError):            # null pointers are rejected before any launch
        _lib.call("sgb_segment_pool_max_fwd", None, 10, 64, None, 10, None, 2, None, None, None, 0, None)
    with pytest.raises(_lib.SgbError):
        _lib.call("sgb_cluster_knn", None, 3, 10, None, None, 1, 20, None, None, 0, None)


def test_ops_refuse_cpu_tensors():
    from seggroup_b200 import _lib, ops
    with pytest.raises(_lib.SgbError):
        ops.exclusive_scan(torch.zeros(4, dtype=torch.int32))


def test_host_label_writer(tmp_path):
    from seggroup_b200 import _lib
    v = np.array([0, -1, 7, 149999, -1, 2147483647], np.int32)
    p = str(tmp_path / "x.txt")
    _lib.call("sgb_write_labels_host", p.encode(), v, len(v))
    assert open(p).read() == "".join("%d\n" % x for x in v)


def test_host_label_writer_every_path(tmp_path):
    """All three formatting paths of sgb_write_labels_host against '%d\\n': the line table ([-1, 2^20 - 2]), the register-assembled
    lines (|v| < 10^8, every digit count, negative values) and the snprintf path beyond; runs re-use the previous line."""
    from seggroup_b200 import _lib
    rng = np.random.default_rng(5)
    edges = [0, -1, -2, 1, 9, 10, 99, 100, 999, 1000, 9999, 10000, 99999, 100000, 999999, 1000000, 1048573, 1048574, 1048575, 1048576,
             9999999, 10000000, 99999999, 100000000, 999999999, 2147483647, -2147483648, -99999999, -100000000, -10000000, -9999999]
    v = np.concatenate([np.arange(-300, 3000), np.array(edges), rng.integers(-2**31, 2**31 - 1, 20000), rng.integers(0, 10**8, 20000),
                        rng.integers(-10**8, 0, 20000), rng.integers(1040000, 1060000, 5000), np.repeat(rng.integers(0, 3000, 500), 7)]).astype(np.int32)
    p = str(tmp_path / "y.txt")
    _lib.call("sgb_write_labels_host", p.encode(), v, len(v))
    assert open(p).read() == "".join("%d\n" % x for x in v)
    _lib.call("sgb_write_labels_host", p.encode(), v[:0], 0)             # empty vector -> empty file
    assert open(p).read() == ""
    with pytest.raises(_lib.SgbError):
        _lib.call("sgb_write_labels_host", str(tmp_path / "no_such_dir" / "z.txt").encode(), v, len(v))


def test_scene_file_loader_roundtrip(tmp_path, scene8k):
    from seggroup_b200 import synth
    from seggroup_b200.model import load_scene_files
    synth.write_scene_tree(str(tmp_path), [scene8k])
    adj, unmap, so, sm = load_scene_files(scene8k.name, os.path.join(str(tmp_path), "dataset", "scannet"))
    assert np.array_equal(so, scene8k.seg_offsets) and np.array_equal(sm, scene8k.seg_members)
    assert np.array_equal(adj, scene8k.adj) and np.array_equal(unmap, scene8k.unmap)
