"""Model check of the flag-in-data all-gather of qr_panel_reg_kernel (csrc/npw_qr_f64.cu), in the spirit of
tests/test_i8emu_protocol.py: the CUDA kernel cannot run here, but its synchronisation argument can be checked.

Protocol (per launch, steps 0..W): every CTA publishes its vector for step s into slot [s & 1][cta] tagged with sequence
number seq0 + s; the first CTA of every group of GSZ polls its members' slots, adds them and publishes the group sum into
group slot [s & 1][group]; the owner of the pivot row publishes it into the pivot slot [s & 1]; every CTA polls all group
slots (and the pivot slot) of step s, then computes and publishes step s + 1.  Packets validate themselves by sequence
number, there is no barrier.  Claim: two parities suffice — a slot is never overwritten (by step s + 2) before every
reader of step s has consumed it, and nobody waits forever — under ANY interleaving of the CTAs."""
import random

import pytest


def simulate(G, W, gsz, seed, seq0=1, max_events=2_000_000):
    rnd = random.Random(seed)
    ngroups = (G + gsz - 1) // gsz
    cta_slot = [[0] * G for _ in range(2)]          # sequence number last written (0 = zeroed scratch)
    grp_slot = [[0] * ngroups for _ in range(2)]
    piv_slot = [0, 0]
    pivot_owner = 0
    # per-CTA program counter: (step, phase); phases: 0 publish own, 1 (leader) gather members, 2 (leader) publish group,
    # 3 gather groups (+ pivot), 4 step done
    step = [0] * G
    phase = [0] * G
    pending = [None] * G                            # set of slots still to be read in the current gather
    consumed = {}                                   # (kind, parity, idx, seq) -> number of readers that consumed it
    done = 0
    events = 0
    while done < G:
        events += 1
        assert events < max_events, "no progress: the protocol deadlocked"
        c = rnd.randrange(G)
        if step[c] > W:
            continue
        s, par = step[c], step[c] & 1
        seq = seq0 + s
        leader, grp = c % gsz == 0, c // gsz
        if phase[c] == 0:
            # publishing step s overwrites the packet of step s - 2: every reader of that packet must be done with it
            old = cta_slot[par][c]
            if old:
                assert consumed.get(("cta", par, c, old), 0) == 1, ("member packet overwritten before its leader read it", c, s)
            cta_slot[par][c] = seq
            if c == pivot_owner and s < W:
                oldp = piv_slot[par]
                if oldp:
                    assert consumed.get(("piv", par, 0, oldp), 0) == G, ("pivot packet overwritten early", s)
                piv_slot[par] = seq
            phase[c] = 1 if leader else 3
            pending[c] = None
        elif phase[c] == 1:
            if pending[c] is None:
                pending[c] = set(range(grp * gsz, min(G, grp * gsz + gsz)))
            for m in list(pending[c]):
                got = cta_slot[par][m]
                assert got <= seq, ("a member ran two steps ahead of its leader", c, m, s)
                if got == seq:
                    consumed[("cta", par, m, seq)] = consumed.get(("cta", par, m, seq), 0) + 1
                    pending[c].discard(m)
            if not pending[c]:
                phase[c] = 2
                pending[c] = None
        elif phase[c] == 2:
            old = grp_slot[par][grp]
            if old:
                assert consumed.get(("grp", par, grp, old), 0) == G, ("group packet overwritten before everybody read it", grp, s)
            grp_slot[par][grp] = seq
            phase[c] = 3
        elif phase[c] == 3:
            if pending[c] is None:
                pending[c] = {("grp", g) for g in range(ngroups)}
                if s < W:
                    pending[c].add(("piv", 0))
            for kind, idx in list(pending[c]):
                got = grp_slot[par][idx] if kind == "grp" else piv_slot[par]
                assert got <= seq, ("a packet of a later step replaced the one this CTA still needs", c, kind, idx, s)
                if got == seq:
                    key = (kind, par, idx, seq)
                    consumed[key] = consumed.get(key, 0) + 1
                    pending[c].discard((kind, idx))
            if not pending[c]:
                pending[c] = None
                phase[c] = 0
                step[c] += 1
                if step[c] > W:
                    done += 1
    return events


@pytest.mark.parametrize("G,gsz", [(1, 12), (3, 12), (12, 12), (13, 12), (37, 4), (148, 12)])
def test_two_parities_suffice_under_random_interleavings(G, gsz):
    for seed in range(6 if G < 100 else 2):
        simulate(G, W=8 if G >= 100 else 33, gsz=gsz, seed=seed)


def test_a_single_parity_would_not_be_enough():
    """The checker itself must be able to fail: with ONE buffer a fast CTA overwrites a packet a slow one still needs."""
    def one_parity(G, W, gsz, seed):
        # same protocol with every slot index forced to parity 0
        g = dict(simulate.__globals__)
        code = compile(open(__file__).read().replace("step[c] & 1", "0").replace("def simulate", "def simulate1"), __file__, "exec")
        exec(code, g)
        return g["simulate1"](G, W, gsz, seed)
    with pytest.raises(AssertionError):
        for seed in range(20):
            one_parity(13, 33, 12, seed)
