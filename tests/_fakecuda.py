"""TEST HARNESS: run the REAL stream engine (numpywren_b200.job_runner.TileEngine.run_node, lambdapack_run) on the host.

The engine is host code: it picks a stream, waits on tile events, decides which input buffer an output may overwrite,
calls a kernel wrapper, stores the result by reference, records an event, reclaims dead tiles.  None of those decisions
needs a GPU to be checked — only the CUDA objects they are expressed with.  ``install(monkeypatch)`` replaces, for one
test, torch.cuda streams / events by inert stand-ins (everything then executes synchronously, in enqueue order, which is
one of the orders the real streams may produce) and the C-ABI by tests/_hostlib.HostLib.  The product code is not
modified and never imports this file; on a GPU box the `-m gpu` tests run the same engine against the CUDA kernels.
"""
import contextlib

import torch

import _hostlib


class FakeEvent:
    def __init__(self, enable_timing=False, **kw):
        self.recorded = False

    def record(self, stream=None):
        self.recorded = True

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 0.0


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **kw):
        self.waited = 0

    def wait_event(self, ev):
        assert isinstance(ev, FakeEvent) and ev.recorded, "waiting on an event that was never recorded"
        self.waited += 1

    def wait_stream(self, other):
        pass

    def synchronize(self):
        pass


class FakePool:
    def __init__(self, n_normal, n_high):
        self.normal = [FakeStream() for _ in range(n_normal)]
        self.high = [FakeStream() for _ in range(n_high)]


def install(monkeypatch):
    """-> HostLib.  After this, BigMatrix(device="cpu") tiles run through lambdapack_run with the real TileEngine."""
    from numpywren_b200 import job_runner
    lib = _hostlib.install(monkeypatch)
    current = FakeStream()
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: current)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, stream: None)
    monkeypatch.setattr(job_runner.StreamPool, "get", classmethod(lambda cls, device, n_normal, n_high: FakePool(n_normal, n_high)))

    def ensure_device(self, tensor_device):
        if self.pool is None:
            self.device = tensor_device
            self.pool = job_runner.StreamPool.get(tensor_device, self.n_streams, self.n_high)
            self._entry_event = FakeEvent()
            self._entry_event.record(current)
            self._entered = set()
    monkeypatch.setattr(job_runner.TileEngine, "_ensure_device", ensure_device)
    return lib
