"""The kernel-seam drop-in of INTEGRATION.md §2 (tools/reference_side_stub/kernels_b200.py): syrk / trsm / chol / qr_factor bound to the
C-ABI with ctypes + the CUDA runtime — NumPy in and out, no torch, nothing imported from numpywren_b200 — driven by the
program logic of algs.CHOLESKY (the oracle's replay, with its three kernels replaced by the stub's) on the fixture written by
the unmodified reference."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import npw_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def stub(cuda_device):
    if cuda_device.type != "cuda":
        pytest.skip("needs a CUDA device (the stub talks to the CUDA runtime directly)")
    os.environ["NPW_B200_LIB"] = os.path.join(ROOT, "numpywren_b200", "lib", "libnpw_b200.so")
    spec = importlib.util.spec_from_file_location("kernels_b200_stub",
                                                  os.path.join(ROOT, "tools", "reference_side_stub", "kernels_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,n,b", [("cholesky_64_16", 64, 16), ("cholesky_60_16", 60, 16)])
def test_reference_program_logic_over_the_stub_matches_the_reference_run(golden_dir, stub, monkeypatch, name, n, b):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    I = orc.OracleBigMatrix("I_stub", (n, n), (b, b))
    orc.shard_matrix(I, g["A"])
    for fn in ("syrk", "trsm", "chol"):
        monkeypatch.setattr(orc, fn, getattr(stub, fn))
    L = orc.run_cholesky(I)[0].numpy()
    ref = np.linalg.cholesky(g["A"])
    assert np.linalg.norm(np.tril(L) - ref) / np.linalg.norm(ref) < 1e-10
    assert np.linalg.norm(np.tril(L) - np.tril(g["L"])) / np.linalg.norm(g["L"]) < 1e-10      # the reference's own factor


def test_stub_chol_reports_non_spd_like_numpy(stub):
    a = np.eye(32)
    a[5, 5] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        stub.chol(a)
    x = np.random.RandomState(0).randn(96, 40)
    spd = x @ x.T + 96 * np.eye(96)
    assert np.allclose(stub.chol(spd), np.linalg.cholesky(spd), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["tsqr_256_32", "tsqr_128_16"])
def test_tsqr_program_logic_over_the_stub_matches_the_reference_run(golden_dir, stub, monkeypatch, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, b, nlev = int(g["m"]), int(g["b"]), int(g["nlev"])
    A = orc.OracleBigMatrix("X_stub", (m, b), (b, b))
    orc.shard_matrix(A, g["X"])
    monkeypatch.setattr(orc, "qr_factor", stub.qr_factor)
    Rs, Vs, Ts = orc.run_tsqr(A)
    R = Rs.get_block(nlev, 0)
    assert np.linalg.norm(R - g["R"]) / np.linalg.norm(g["R"]) < 1e-10
    for k in g.files:
        if k[:2] in ("R_", "V_") or k.startswith("Tq_"):
            lvl, j = (int(x) for x in k.split("_")[1:])
            mat = {"R": Rs, "V": Vs, "Tq": Ts}[k.split("_")[0]]
            got = mat.get_block(lvl, j)
            assert np.linalg.norm(got - g[k]) / max(np.linalg.norm(g[k]), 1e-300) < 1e-10, k
