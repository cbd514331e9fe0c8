"""Round trip through the reference's object layout (np.save tiles + JSON header with base64-pickled dtype)."""
import json
import os

import numpy as np
import pytest

from numpywren_b200 import s3_format
from numpywren_b200.matrix import BigMatrix
from numpywren_b200.matrix_init import shard_matrix
from numpywren_b200.matrix_utils import constant_zeros


def test_export_layout_and_roundtrip(tmp_path, unique_key):
    key = unique_key("fmt")
    X = np.random.RandomState(0).randn(10, 7)
    m = BigMatrix(key, shape=X.shape, shard_sizes=(4, 4), device="cpu")
    m.free()
    shard_matrix(m, X)
    assert s3_format.export_matrix(m, str(tmp_path)) == 6
    base = tmp_path / "numpywren.objects" / key
    names = sorted(os.listdir(base))
    # object names exactly as reference matrix.py:457-464 builds them: "{start}_{end}_{shard}_" per axis
    assert names == sorted(["header", "0_4_4_0_4_4_", "0_4_4_4_7_4_", "4_8_4_0_4_4_", "4_8_4_4_7_4_", "8_10_4_0_4_4_",
                            "8_10_4_4_7_4_"])
    hdr = json.loads((base / "header").read_text())
    assert hdr["shape"] == [10, 7] and hdr["shard_sizes"] == [4, 4]
    assert s3_format.decode_dtype(hdr["dtype"]) == np.float64
    assert np.array_equal(np.load(base / "8_10_4_4_7_4_"), X[8:10, 4:7])          # plain np.save payload
    m.delete()
    back = s3_format.import_matrix(key, str(tmp_path), device="cpu")
    assert back.shape == (10, 7) and back.shard_sizes == (4, 4)
    assert np.array_equal(back.numpy(), X)


def test_partial_matrix_and_3d(tmp_path, unique_key):
    key = unique_key("fmt3")
    s = BigMatrix(key, shape=(3, 8, 8), shard_sizes=(1, 4, 4), device="cpu", parent_fn=constant_zeros)
    s.free()
    s.put_block(np.full((4, 4), 2.0), 1, 0, 1)               # autosqueezed put, stored as (1, 4, 4)
    assert s3_format.export_matrix(s, str(tmp_path)) == 1
    assert np.load(tmp_path / "numpywren.objects" / key / "1_2_1_0_4_4_4_8_4_").shape == (1, 4, 4)
    s.delete()
    back = s3_format.import_matrix(key, str(tmp_path), device="cpu", parent_fn=constant_zeros)
    assert back.block_idxs_exist == [(1, 0, 1)]
    assert float(back.get_block(1, 0, 1).sum()) == 32.0 and not back.get_block(0, 0, 0).any()


def test_header_dtype_accepts_numpy_dtypes_only():
    """The header's dtype is a pickle (reference matrix.py:547-555); import_matrix reads it from a synced directory, so
    only NumPy dtypes may be resolved — a crafted header must not be able to call anything."""
    import base64
    import pickle

    import pytest
    for dt in (np.float64, np.float32, np.int64, np.dtype("float64"), np.dtype("<i4"), np.complex128):
        assert s3_format.decode_dtype(s3_format.encode_dtype(dt)) == dt

    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned > /dev/null",))
    for payload in (pickle.dumps(Evil()), pickle.dumps(os.getcwd), pickle.dumps({"a": 1}), pickle.dumps(np.zeros(2))):
        with pytest.raises(Exception) as ei:
            s3_format.decode_dtype(base64.b64encode(payload).decode())
        assert "dtype" in str(ei.value).lower() or "numpy" in str(ei.value).lower()


def _golden_s3(golden_dir):
    import base64
    d = json.load(open(os.path.join(golden_dir, "s3_format.json")))
    for case in d.values():
        case["X"] = np.frombuffer(base64.b64decode(case["data"]), dtype=np.dtype(case["dtype"])).reshape(case["shape"])
        case["objects"] = {k: base64.b64decode(v) for k, v in case["objects"].items()}
    return d


def _export_and_compare(case, key, root, device):
    """Export ``case`` through this repo and compare every object, name for name and byte for byte, with what the
    unmodified reference sent to S3 for the same matrix (oracle/make_golden.py golden_s3_format)."""
    X = case["X"]
    m = BigMatrix(key, shape=tuple(case["shape"]), shard_sizes=tuple(case["shard_sizes"]), dtype=np.dtype(case["dtype"]).type,
                  device=device)
    m.free()
    shard_matrix(m, X)
    n = s3_format.export_matrix(m, root)
    want = case["objects"]
    assert n == len(want) - 1
    got = {}
    base = os.path.join(root, m.key_base)
    for name in os.listdir(base):
        got[os.path.join(m.key_base, name)] = open(os.path.join(base, name), "rb").read()
    assert sorted(got) == sorted(want)                                   # the reference's own object names
    for k in want:
        if k.endswith("/header"):
            assert json.loads(got[k]) == json.loads(want[k])             # same JSON document ...
            assert s3_format.decode_dtype(json.loads(want[k])["dtype"]) == np.dtype(case["dtype"]).type
        else:
            assert got[k] == want[k], k                                  # ... and the same np.save bytes per tile
    return m


@pytest.mark.parametrize("key", ["fmt2d", "fmt3d", "fmtf32"])
def test_export_equals_the_objects_the_reference_writes(tmp_path, golden_dir, key):
    case = _golden_s3(golden_dir)[key]
    _export_and_compare(case, key, str(tmp_path), "cpu").delete()
    # and the reference's objects, laid out as a synced directory, import to the same matrix
    root = tmp_path / "ref"
    for k, body in case["objects"].items():
        p = root / k
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(body)
    back = s3_format.import_matrix(key, str(root), device="cpu")
    assert back.dtype == np.dtype(case["dtype"]).type and tuple(back.shape) == tuple(case["shape"])
    assert np.array_equal(back.numpy(), case["X"])
    back.delete()


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["fmt2d", "fmt3d"])
def test_export_import_round_trip_from_hbm(tmp_path, golden_dir, key, cuda_device):
    """Tiles resident in HBM -> the reference's objects (byte-equal to the reference's own) -> back into HBM."""
    case = _golden_s3(golden_dir)[key]
    m = _export_and_compare(case, key, str(tmp_path), cuda_device)
    m.delete()
    back = s3_format.import_matrix(key, str(tmp_path), device=cuda_device)
    assert all(back._get_block_ref(*b).is_cuda for b in back.block_idxs_exist)
    assert np.array_equal(back.numpy(), case["X"])
    back.delete()
