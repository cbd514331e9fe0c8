"""BigMatrix storage semantics on a storage-only host device (no kernels run): the reference's
tests/test_simple.py, test_indexing.py, test_transpose.py, test_multiaxis.py restated, plus parent_fn /
lambdav / autosqueeze / safe / header behaviour from matrix.py."""
import numpy as np
import pytest
import torch

from numpywren_b200 import matrix_utils
from numpywren_b200.matrix import BigMatrix, BigMatrixView
from numpywren_b200.matrix_init import local_numpy_init, shard_matrix
from numpywren_b200.utils import convert_to_slice


def cpu_matrix(key, shape, shard_sizes, **kw):
    m = BigMatrix(key, shape=shape, shard_sizes=shard_sizes, device="cpu", **kw)
    m.free()
    return m


def test_single_shard_matrix(unique_key):
    # tests/test_simple.py:12-18
    X = np.random.randn(128, 128)
    X_sharded = cpu_matrix(unique_key(), X.shape, X.shape)
    shard_matrix(X_sharded, X)
    assert np.all(X_sharded.numpy() == X)


def test_multiple_shard_matrix_uneven(unique_key):
    # tests/test_simple.py:30-45: 200 with shards of 101, then reload through the header
    X = np.random.randn(200, 200)
    key = unique_key()
    X_sharded = cpu_matrix(key, X.shape, (101, 101), write_header=True)
    shard_matrix(X_sharded, X)
    assert X_sharded.num_blocks(0) == 2 and X_sharded.get_block(1, 1).shape == (99, 99)
    assert np.all(X_sharded.numpy() == X)
    again = BigMatrix(key, device="cpu")
    assert again.shape == (200, 200) and again.shard_sizes == (101, 101)
    assert np.all(again.numpy() == X)
    again.delete()
    with pytest.raises(Exception, match="Header doesn't exist"):
        BigMatrix(key, device="cpu")


def test_get_block_returns_private_copy(unique_key):
    m = cpu_matrix(unique_key(), (8, 8), (4, 4))
    shard_matrix(m, np.ones((8, 8)))
    t = m.get_block(0, 0)
    t += 5
    assert float(m.get_block(0, 0).sum()) == 16.0
    src = torch.ones(4, 4, dtype=torch.float64)
    m.put_block(src, 1, 1)
    src += 1
    assert float(m.get_block(1, 1).sum()) == 16.0


def test_missing_block_parent_fn_and_errors(unique_key):
    m = cpu_matrix(unique_key(), (10, 10), (4, 4))
    with pytest.raises(Exception, match="not exist"):
        m.get_block(0, 0)
    with pytest.raises(Exception, match="does not match shape"):
        m.get_block(0)
    z = cpu_matrix(unique_key(), (10, 10), (4, 4), parent_fn=matrix_utils.constant_zeros)
    assert z.get_block(2, 2).shape == (2, 2) and not z.get_block(2, 0).any()
    assert not z.numpy().any()
    c = cpu_matrix(unique_key(), (10, 10), (4, 4), parent_fn=matrix_utils.make_constant_parent(7.0))
    assert float(c.get_block(0, 2).sum()) == 7.0 * 8
    with pytest.raises(Exception, match="same length"):
        BigMatrix(unique_key(), shape=(4, 4), shard_sizes=(2,), device="cpu")


def test_lambdav_only_on_diagonal_tiles_of_square_matrices(unique_key):
    # matrix.py:307-309 and :129-130
    m = cpu_matrix(unique_key(), (8, 8), (4, 4), lambdav=2.5)
    shard_matrix(m, np.zeros((8, 8)))
    assert np.array_equal(m.get_block(1, 1).numpy(), 2.5 * np.eye(4))
    assert not m.get_block(1, 0).any()
    assert np.array_equal(m.get_block(1, 1).numpy(), 2.5 * np.eye(4))   # stored tile untouched: shift is on read
    with pytest.raises(Exception, match="square"):
        BigMatrix(unique_key(), shape=(8, 4), shard_sizes=(4, 4), lambdav=1.0, device="cpu")


def test_autosqueeze_and_safe(unique_key):
    s = cpu_matrix(unique_key(), (3, 8, 8), (1, 4, 4), parent_fn=matrix_utils.constant_zeros)
    assert s.get_block(0, 1, 1).shape == (4, 4)                  # squeezed on read
    s.put_block(np.ones((4, 4)), 2, 0, 1)                          # re-expanded on write
    assert s._get_block_ref(2, 0, 1).shape == (1, 4, 4)
    with pytest.raises(Exception, match="Incompatible block size"):
        s.put_block(np.ones((3, 3)), 2, 0, 1)
    u = cpu_matrix(unique_key(), (3, 8, 8), (1, 4, 4), safe=False)
    u.put_block(np.ones((3, 3)), 2, 0, 1)                          # TSQR relies on safe=False (alg_wrappers.py:36)
    ns = cpu_matrix(unique_key(), (3, 8, 8), (1, 4, 4), autosqueeze=False, parent_fn=matrix_utils.constant_zeros)
    assert ns.get_block(0, 1, 1).shape == (1, 4, 4)


def test_sharded_matrix_row_get(unique_key):
    # tests/test_indexing.py:11-20
    X = np.random.randn(128, 128)
    m = cpu_matrix(unique_key(), X.shape, (32, 32))
    shard_matrix(m, X)
    sub = m.submatrix(0)
    assert np.all(sub.numpy() == X[0:32])
    sub = m.submatrix(None, 1)
    assert np.all(sub.numpy() == X[:, 32:64])


def test_complex_slices(unique_key):
    # tests/test_indexing.py:22-40
    X = np.random.randn(21, 67, 53)
    m = cpu_matrix(unique_key(), X.shape, (21, 16, 11))
    shard_matrix(m, X)
    assert np.all(m.submatrix(0, [2, 4]).numpy() == X[:, 32:64])
    assert np.all(m.submatrix(0, [1, None, 2], [1, None]).numpy()[:, :16] == X[:, 16:32, 11:])


def test_step_slices(unique_key):
    # tests/test_indexing.py:42-60
    X = np.random.randn(128, 128)
    m = cpu_matrix(unique_key(), X.shape, (16, 16))
    shard_matrix(m, X)
    got = m.submatrix([None, None, 2]).numpy()
    want = np.vstack([X[i:i + 16] for i in range(0, 128, 32)])
    assert np.all(got == want)
    got = m.submatrix(None, [1, 6, 3]).numpy()
    want = np.hstack([X[:, 16:32], X[:, 64:80]])
    assert np.all(got == want)


def test_transpose_views(unique_key):
    # tests/test_transpose.py:11-27
    X = np.random.randn(20, 33)
    m = cpu_matrix(unique_key(), X.shape, (10, 11))
    shard_matrix(m, X)
    assert m.T.shape == (33, 20) and m.T.shard_sizes == (11, 10)
    assert np.all(m.T.numpy() == X.T)
    assert np.all(m.T.get_block(2, 1).numpy() == X.T[22:33, 10:20])
    assert m.T.true_block_idx(2, 1) == (1, 2)
    t = cpu_matrix(unique_key(), X.T.shape, (11, 10))
    shard_matrix(t.T, X)                                   # writes go through the transposed view
    assert np.all(t.numpy() == X.T)


def test_multiaxis(unique_key):
    # tests/test_multiaxis.py:11-29
    X = np.random.randn(8, 8, 8, 8)
    m = cpu_matrix(unique_key(), X.shape, (4, 4, 4, 4))
    shard_matrix(m, X)
    assert np.all(m.numpy() == X)
    X3 = np.random.randn(21, 67, 53)
    m3 = cpu_matrix(unique_key(), X3.shape, (21, 16, 11))
    shard_matrix(m3, X3)
    assert np.all(m3.numpy() == X3)
    assert m3.get_block(0, 4, 4).shape == (21, 3, 9)


def test_block_bookkeeping(unique_key):
    m = cpu_matrix(unique_key(), (6, 6), (4, 4), write_header=True)
    assert m.block_idxs == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert m.blocks[3] == ((4, 6), (4, 6))
    assert m.block_idxs_exist == [] and len(m.block_idxs_not_exist) == 4
    m.put_block(np.zeros((4, 2)), 0, 1)
    assert m.block_idxs_exist == [(0, 1)] and (0, 1) not in m.block_idxs_not_exist
    assert m.__shard_idx_to_key__((0, 1)).endswith("0_4_4_4_6_4_")     # reference object naming (matrix.py:457-464)
    m.delete_block(0, 1)
    assert m.block_idxs_exist == []
    m.put_block(np.zeros((4, 4)), 0, 0)
    assert m.free() == 0 and m.block_idxs_exist == []
    assert isinstance(m.submatrix(0), BigMatrixView)


def test_same_key_shares_storage(unique_key):
    key = unique_key()
    a = cpu_matrix(key, (4, 4), (2, 2))
    b = BigMatrix(key, shape=(4, 4), shard_sizes=(2, 2), device="cpu")
    a.put_block(np.full((2, 2), 3.0), 1, 0)
    assert float(b.get_block(1, 0).sum()) == 12.0


def test_local_numpy_init(unique_key):
    X = np.random.randn(10, 6)
    m = local_numpy_init(X, (4, 4), device="cpu")
    assert np.all(m.numpy() == X)


def test_convert_to_slice():
    assert convert_to_slice(None) == slice(None, None, None)
    assert convert_to_slice(3) == slice(3, 4, 1)
    assert convert_to_slice([5]) == slice(None, 5, None)
    assert convert_to_slice([1, 5]) == slice(1, 5, None)
    assert convert_to_slice([1, 5, 2]) == slice(1, 5, 2)
    with pytest.raises(ValueError):
        convert_to_slice([1, 2, 3, 4])


# ------------------------------------------------------------------ reshard_down (reference tests/test_reshard.py:12-48)
@pytest.mark.parametrize("shape,shards,breaks", [((128, 128), (128, 128), (4, 4)), ((128, 128), (64, 64), (4, 4)),
                                                  ((100, 72), (64, 48), (2, 3))])
def test_reshard_down_matrix(unique_key, shape, shards, breaks):
    from numpywren_b200.matrix_init import reshard_down
    X = np.random.RandomState(0).randn(*shape)
    A = BigMatrix(unique_key("rs"), shape=shape, shard_sizes=shards, device="cpu")
    shard_matrix(A, X)
    B = reshard_down(A, breaks)
    assert tuple(B.shard_sizes) == tuple(s // k for s, k in zip(shards, breaks))
    assert np.all(A.numpy() == X) and np.all(B.numpy() == X)
    assert B.get_block(0, 0).shape == tuple(min(s // k, n) for s, k, n in zip(shards, breaks, shape))
    A.free(); B.free()


def test_reshard_down_tensor(unique_key):
    from numpywren_b200.matrix_init import reshard_down
    X = np.random.RandomState(1).randn(128, 128, 4)
    A = BigMatrix(unique_key("rs3"), shape=X.shape, shard_sizes=(64, 64, 4), device="cpu")
    A.autosqueeze = False
    shard_matrix(A, X)
    B = reshard_down(A, (4, 4, 2), pwex=None)
    assert np.all(A.numpy() == X) and np.all(B.numpy() == X)
    assert tuple(B.get_block(0, 0, 0).shape) == (16, 16, 2)
    A.free(); B.free()
