"""Protocol model of csrc/npw_ozaki_i8.cu's producer / MMA / epilogue synchronisation (CPU, no CUDA).

The kernel's mbarrier usage is transcribed literally — same barriers, same wait parities, same points at which
tcgen05.commit and TMA completions arrive — and run under randomised interleavings with asynchronous completion of TMA
loads and MMAs.  Checked: no deadlock; an X slot / Y buffer is never overwritten while an MMA that reads it is still
in flight; no MMA reads a slot whose TMA load has not landed (or holds a different k-block / digit); every (k-block,
digit pair) is multiplied exactly once; the epilogue starts only after the last MMA completed.
A hang on the GPU costs a box; a parity slip here costs nothing."""
import random

import pytest

NX = 4


class MBar:
    """mbarrier with arrival count 1 (+ optional transaction bytes): `completed` counts finished phases."""

    def __init__(self):
        self.completed = 0
        self.pending_tx = 0
        self.armed = False

    def try_wait(self, parity):            # mbarrier.try_wait.parity: true iff the phase with this parity has completed
        return (self.completed % 2) != parity

    def arrive(self):                      # plain arrival (tcgen05.commit)
        self.completed += 1

    def arrive_expect_tx(self, nbytes):    # producer's arrival + expected bytes; phase completes when the bytes landed
        self.armed, self.pending_tx = True, self.pending_tx + nbytes

    def complete_tx(self, nbytes):
        self.pending_tx -= nbytes
        if self.armed and self.pending_tx == 0:
            self.armed = False
            self.completed += 1


def run(s, kblocks, seed):
    rnd = random.Random(seed)
    x_full, x_empty = [MBar() for _ in range(NX)], [MBar() for _ in range(NX)]
    y_full, y_empty = [MBar(), MBar()], [MBar(), MBar()]
    acc_full = MBar()
    xslot = [None] * NX                    # (kb, digit) landed in the slot, None while a load is in flight
    ybuf = [None, None]                    # kb whose Y digits landed
    tma_q = []                             # in-flight TMA loads: (kind, index, payload, bar, bytes)
    mma_q = []                             # issued, not yet completed MMAs / commits, in issue order
    readers = {("x", i): 0 for i in range(NX)}
    readers.update({("y", i): 0 for i in range(2)})
    done_pairs = set()
    state = {"epilogue": False, "mma_done": False}

    def x_producer():
        xit = 0
        for kb in range(kblocks):
            for pd in range(s):
                slot = xit % NX
                while not x_empty[slot].try_wait(((xit // NX) & 1) ^ 1):
                    yield
                assert readers[("x", slot)] == 0, "X slot overwritten while an MMA still reads it"
                x_full[slot].arrive_expect_tx(1)
                xslot[slot] = None
                tma_q.append(("x", slot, (kb, pd), x_full[slot], 1))
                xit += 1
                yield

    def y_producer():
        for kb in range(kblocks):
            yb = kb & 1
            while not y_empty[yb].try_wait(((kb >> 1) & 1) ^ 1):
                yield
            assert readers[("y", yb)] == 0, "Y buffer overwritten while an MMA still reads it"
            y_full[yb].arrive_expect_tx(s)
            ybuf[yb] = None
            for q in range(s):
                tma_q.append(("y", yb, (kb, q), y_full[yb], 1))
            yield

    def mma_issuer():
        xit = 0
        for kb in range(kblocks):
            yb = kb & 1
            while not y_full[yb].try_wait((kb >> 1) & 1):
                yield
            for pd in range(s):
                slot = xit % NX
                while not x_full[slot].try_wait((xit // NX) & 1):
                    yield
                assert xslot[slot] == (kb, pd), f"MMA would read slot {slot} = {xslot[slot]}, wants {(kb, pd)}"
                assert ybuf[yb] == kb, f"MMA would read Y buffer {yb} = {ybuf[yb]}, wants {kb}"
                for q in range(s - pd):
                    readers[("x", slot)] += 1
                    readers[("y", yb)] += 1
                    mma_q.append(("mma", slot, yb, (kb, pd, q)))
                mma_q.append(("commit", x_empty[slot]))
                xit += 1
                yield
            mma_q.append(("commit", y_empty[yb]))
        mma_q.append(("commit", acc_full))
        state["mma_done"] = True

    def epilogue():
        while not acc_full.try_wait(0):
            yield
        assert state["mma_done"] and not any(e[0] == "mma" for e in mma_q), "epilogue before the last MMA completed"
        state["epilogue"] = True

    agents = [x_producer(), y_producer(), mma_issuer(), epilogue()]
    alive = [True] * len(agents)
    idle_rounds = 0
    while any(alive):
        progressed = False
        order = list(range(len(agents)))
        rnd.shuffle(order)
        for a in order:
            if alive[a] and rnd.random() < 0.7:
                before = (len(tma_q), len(mma_q))
                try:
                    next(agents[a])
                except StopIteration:
                    alive[a] = False
                    progressed = True
                if (len(tma_q), len(mma_q)) != before:
                    progressed = True
        # asynchronous hardware: TMA loads land in any order, MMAs / commits complete in issue order
        if tma_q and rnd.random() < 0.6:
            kind, idx, payload, bar, nbytes = tma_q.pop(rnd.randrange(len(tma_q)))
            if kind == "x":
                xslot[idx] = payload
            else:
                ybuf[idx] = payload[0]
            bar.complete_tx(nbytes)
            progressed = True
        if mma_q and rnd.random() < 0.6:
            ev = mma_q.pop(0)
            if ev[0] == "mma":
                _, slot, yb, pair = ev
                readers[("x", slot)] -= 1
                readers[("y", yb)] -= 1
                assert pair not in done_pairs
                done_pairs.add(pair)
            else:
                ev[1].arrive()
            progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 2000, "deadlock: no agent can make progress"
    assert state["epilogue"]
    assert done_pairs == {(kb, pd, q) for kb in range(kblocks) for pd in range(s) for q in range(s - pd)}


@pytest.mark.parametrize("s", [1, 2, 3, 6, 8])
@pytest.mark.parametrize("kblocks", [1, 2, 3, 7, 32])
def test_barrier_protocol_has_no_deadlock_and_no_hazard(s, kblocks):
    for seed in range(6):
        run(s, kblocks, seed)


def test_hand_packed_umma_descriptors_equal_cutlass_bitfields(tmp_path):
    """csrc/npw_ozaki_i8.cu packs the tcgen05 shared-memory / instruction descriptors by hand; the vendored CUTLASS headers
    define the same fields as bitfield structs.  Compile a host program that builds both and compare."""
    import os
    import shutil
    import site
    import subprocess
    inc = None
    for sp in site.getsitepackages():
        cand = os.path.join(sp, "flashinfer", "data", "cutlass", "include")
        if os.path.exists(os.path.join(cand, "cute", "arch", "mma_sm100_desc.hpp")):
            inc = cand
    if inc is None or shutil.which("nvcc") is None:
        pytest.skip("CUTLASS headers or nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "umma_desc_check")
    subprocess.run(["nvcc", "-std=c++17", "-w", "-I" + inc, "-o", exe, os.path.join(root, "tools", "umma_desc_check.cu")],
                   check=True, capture_output=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout
    # and the kernel source really uses the same packing as the checked copy
    src = open(os.path.join(root, "numpywren_b200", "csrc", "npw_ozaki_i8.cu")).read()
    for frag in ("<< 16;", "(1024 >> 4) << 32;", "<< 46;", "<< 61;", "(2u << 4) | (1u << 7) | (1u << 10)"):
        assert frag in src
