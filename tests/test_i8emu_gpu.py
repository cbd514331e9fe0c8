"""fp64 syrk emulated on the int8 tensor cores (csrc/npw_ozaki_i8.cu, tcgen05.mma kind::i8): digit extraction and the
product kernel against the CPU prototype of the same arithmetic (tools/ozaki_prototype.py), and a whole Cholesky with the
emulated syrk against the oracle at the parity bar.  First ran on a B200 in round 2 (profiles/r02a_tcgen05_i8_probe.log,
r02c_syrk_i8emu_timing.jsonl); the engine uses the kernel only with NPW_B200_SYRK=i8emu."""
import os
import sys

import numpy as np
import pytest
import torch

from numpywren_b200 import kernels

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import ozaki_prototype as oz  # noqa: E402

pytestmark = pytest.mark.gpu


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.mark.parametrize("digits", [1, 6, 8])
def test_split_i8_matches_the_prototype_bit_for_bit(cuda_device, digits):
    rs = np.random.RandomState(digits)
    x = rs.randn(256, 384) * np.exp(rs.uniform(-20, 20, size=256))[:, None]
    x[3] = 0.0
    x[5, 0] = 2.0 ** 10
    d, e = kernels.split_i8(dev(x, cuda_device), digits)
    dref, eref, _ = oz.split_rows(x, digits)
    assert np.array_equal(e.cpu().numpy(), eref.astype(np.int32))
    assert np.array_equal(d.cpu().numpy(), dref)


@pytest.mark.parametrize("m,n,k,digits", [(128, 64, 128, 1), (128, 64, 128, 6), (128, 64, 512, 8), (256, 192, 1024, 6),
                                          (4096, 4096, 4096, 6)])
def test_syrk_i8emu_matches_the_prototype(cuda_device, m, n, k, digits):
    rs = np.random.RandomState(m + n + k + digits)
    x, y, s = rs.randn(m, k), rs.randn(n, k), rs.randn(m, n)
    xd, xe = kernels.split_i8(dev(x, cuda_device), digits)
    yd, ye = kernels.split_i8(dev(y, cuda_device), digits)
    c = kernels.syrk_i8emu(dev(s, cuda_device), xd, xe, yd, ye).cpu().numpy()
    if m <= 256:
        ref = s - oz.ozaki_gemm_nt(x, y, digits)[0]           # same digits, same dropped pairs: agreement to rounding
        assert np.abs(c - ref).max() <= 1e-13 * np.abs(ref).max()
    exact = s - x @ y.T
    tol = {1: 0.3, 6: 1e-10, 8: 1e-13}[digits]
    assert np.linalg.norm(c - exact) / np.linalg.norm(exact) < tol


def test_syrk_i8emu_in_place_and_lower_only(cuda_device):
    rs = np.random.RandomState(1)
    x, s = rs.randn(512, 256), rs.randn(512, 512)
    xd, xe = kernels.split_i8(dev(x, cuda_device), 6)
    st = dev(s, cuda_device)
    out = kernels.syrk_i8emu(st, xd, xe, xd, xe, out=st, lower=True)
    assert out.data_ptr() == st.data_ptr()
    c = out.cpu().numpy()
    exact = s - x @ x.T
    i, j = np.indices(c.shape)
    low = j // 64 * 64 <= i // 128 * 128 + 127              # 128 x 64 tiles touching the lower triangle are computed
    # 6 digits: 4.7e-12 of |x_i||y_j| per entry (profiles/r02b_syrk_i8emu_timing.jsonl), |x_i|^2 ~ k = 256
    assert np.abs(c[low] - exact[low]).max() < 1e-10 * 256
    assert np.array_equal(c[~low], s[~low])                  # the others are left alone


@pytest.mark.parametrize("digits", [7, 8])
def test_cholesky_with_emulated_syrk_against_oracle(unique_key, cuda_device, monkeypatch, digits):
    """N=2048 with 256-tiles, every syrk product on the int8 tensor cores (trsm / potrf native fp64): the factor still
    matches the oracle's at the 1e-10 parity bar (7 digits = 48 mantissa bits per row-scaled entry, 8 digits = 55)."""
    from numpywren_b200 import job_runner
    from numpywren_b200 import lambdapack as lp
    from numpywren_b200.alg_wrappers import cholesky
    from numpywren_b200.matrix import BigMatrix
    from oracle import npw_oracle as orc
    monkeypatch.setenv("NPW_B200_SYRK", "i8emu")
    monkeypatch.setenv("NPW_B200_I8_DIGITS", str(digits))
    n, b = 2048, 256
    nb = n // b
    A = BigMatrix(unique_key(f"i8chol{digits}"), shape=(n, n), shard_sizes=(b, b))
    I = orc.OracleBigMatrix("I", (n, n), (b, b))
    for j in range(nb):
        for k in range(nb):
            t = orc.spd_tile(j, k, b, n, width=64)
            I.put_block(t, j, k)
            A.put_block(t, j, k)
    O_ref, _ = orc.run_cholesky(I)
    program, meta = cholesky(A)
    before = kernels._capi.launch_count()
    program.start()
    job_runner.lambdapack_run(program, timeout=120)
    assert program.program_status() == lp.PS.SUCCESS
    assert program._engine.syrk_mode == "i8emu"
    L, Lref = meta["outputs"][0].numpy(), O_ref.numpy()
    assert np.linalg.norm(L - Lref) / np.linalg.norm(Lref) < 1e-10
    assert kernels._capi.launch_count() > before
