"""CPU model of the producer / MMA-issuer hand-shake of k_ozaki_gemm (csrc/ozaki.cu): the slice-unit ring,
the mbarrier parities and the order in which slots are handed back.  Random completion delays; checks that
the schedule neither deadlocks nor reads a slot that holds the wrong unit.  No GPU."""

import random

import pytest


def oz_order(idx, S):
    return S - (idx >> 1) if (idx & 1) else 1 + (idx >> 1)


def oz_pos(p, S):
    return 2 * (p - 1) if 2 * p <= S + 1 else 2 * (S - p) + 1


class MBar(object):
    """mbarrier with arrival count 1: `done` = number of completed phases; wait(parity) passes once the phase
    of that parity has completed, i.e. when the barrier's current phase parity differs from it."""

    def __init__(self):
        self.done = 0

    def arrive(self):
        self.done += 1

    def test(self, parity):
        return (self.done & 1) != parity


def ring_slots(S, max_slots):
    return min(2 * S + 4, max_slots)


def simulate(S, KB, seed, max_slots=18):
    R = ring_slots(S, max_slots)
    rng = random.Random(seed)
    full = [MBar() for _ in range(R)]
    empty = [MBar() for _ in range(R)]
    slot_content = [None] * R  # unit id whose data is (or is being) in the slot
    slot_ready = [False] * R
    pending = []  # (time, kind, slot, unit)
    now = 0
    # producer state
    pu = 0
    n_units = KB * S
    # mma state: list of steps generated lazily
    def mma_program():
        for kb in range(KB):
            ub = kb * S
            for idx in range(S):
                u = ub + idx
                yield ('wait_full', u % R, (u // R) & 1, u)
            r = 1
            while 2 * r <= S + 1:
                for t in range(r, S + 2 - r):
                    for side in (0, 1):
                        if side == 1 and t == r:
                            continue
                        pa, pb = (r, t) if side == 0 else (t, r)
                        assert 2 <= pa + pb <= S + 1
                        yield ('mma', (ub + oz_pos(pa, S)) % R, ub + oz_pos(pa, S), (ub + oz_pos(pb, S)) % R, ub + oz_pos(pb, S))
                yield ('commit', (ub + oz_pos(r, S)) % R)
                if S + 1 - r != r:
                    yield ('commit', (ub + oz_pos(S + 1 - r, S)) % R)
                r += 1
        yield ('done',)

    prog = mma_program()
    cur = next(prog)
    pairs = 0
    outstanding_mma = 0  # MMAs issued whose completion has not happened (they complete in order)
    mma_done_at = 0
    steps = 0
    while True:
        steps += 1
        assert steps < 200000, 'deadlock'
        progressed = False
        # completions
        for ev in sorted([e for e in pending if e[0] <= now]):
            pending.remove(ev)
            _, kind, slot, unit = ev
            if kind == 'tma':
                assert slot_content[slot] == unit
                slot_ready[slot] = True
                full[slot].arrive()
            else:  # commit completion: the slot is handed back
                empty[slot].arrive()
            progressed = True
        # producer
        if pu < n_units:
            slot, rnd = pu % R, pu // R
            if rnd == 0 or empty[slot].test((rnd - 1) & 1):
                assert oz_order(pu % S, S) in range(1, S + 1)
                slot_content[slot] = pu
                slot_ready[slot] = False
                pending.append((now + rng.randint(1, 30), 'tma', slot, pu))
                pu += 1
                progressed = True
        # mma issuer
        if cur[0] == 'wait_full':
            _, slot, parity, unit = cur
            if full[slot].test(parity):
                assert slot_content[slot] == unit and slot_ready[slot], 'full barrier passed for the wrong unit'
                cur = next(prog)
                progressed = True
        elif cur[0] == 'mma':
            _, sa, ua, sb, ub_ = cur
            assert slot_content[sa] == ua and slot_ready[sa], 'A operand slot overwritten'
            assert slot_content[sb] == ub_ and slot_ready[sb], 'B operand slot overwritten'
            mma_done_at = max(mma_done_at, now) + rng.randint(1, 4)
            pairs += 1
            cur = next(prog)
            progressed = True
        elif cur[0] == 'commit':
            pending.append((max(mma_done_at, now) + 1, 'commit', cur[1], None))
            cur = next(prog)
            progressed = True
        elif cur[0] == 'done':
            if pu == n_units and not pending:
                break
        if not progressed:
            now += 1
    return pairs


@pytest.mark.parametrize('max_slots', [18, 9])  # unit width 64 B (default) / 128 B
@pytest.mark.parametrize('S', [2, 3, 4, 5, 6, 7])
def test_ring_protocol(S, max_slots):
    for KB in (1, 2, 3, 8, 16):
        for seed in range(5):
            pairs = simulate(S, KB, seed, max_slots)
            n_pairs = sum(1 for p in range(1, S + 1) for q in range(1, S + 1) if p + q <= S + 1)
            assert pairs == KB * n_pairs


def test_order_tables():
    for S in range(2, 8):
        order = [oz_order(i, S) for i in range(S)]
        assert sorted(order) == list(range(1, S + 1))
        assert all(oz_pos(order[i], S) == i for i in range(S))
