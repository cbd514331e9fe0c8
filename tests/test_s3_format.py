"""Round trip through the reference's object layout (np.save tiles + JSON header with base64-pickled dtype)."""
import json
import os

import numpy as np

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
