"""numpywren_b200 — the LambdaPACK tile-DAG hot path of numpywren, rebuilt for NVIDIA B200.

Program surface kept from the reference (numpywren/): ``BigMatrix`` get_block/put_block,
the ``algs`` DSL programs, ``alg_wrappers.cholesky/gemm/tsqr/qr/bdfac``, ``lambdapack.LambdaPackProgram``
and ``job_runner.lambdapack_run``.  Underneath: tiles live in HBM, tasks are issued onto CUDA
streams by a host DAG scheduler, and every tile op is a hand-written sm_100a kernel reached
through the C-ABI in ``include/npw_b200.h``.
"""
__version__ = "0.1.0"
