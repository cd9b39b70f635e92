"""CPU model of the one-barrier tile loop of k_predict_main (PCfg OB = 1, csrc/predict.cu): 8 warps with random
progress, the CTA-wide barrier, the two X/JA stages filled by bulk copies (issued by one thread right after the barrier
of tile t for tile t + 1) and the double-buffered C1/C2 tiles.  Checks the two reuse hazards the design rests on:
a stage is never overwritten while a warp still reads it, and a C buffer is never rewritten while a warp still reads it
-- and that the loop cannot deadlock.  No GPU."""

import random

import pytest


class Hazard(Exception):
    pass


def run(n_tiles, n_warps, seed, max_steps=200000):
    rng = random.Random(seed)
    # per warp: (tile, phase); phases: 0 wait stage full, 1 GEMM1 + transform (reads stage, writes C), 2 at barrier,
    # 3 GEMM2 (reads C and stage), then next tile
    tile = [0] * n_warps
    phase = [0] * n_warps
    stage_full = [None, None]  # which tile's data the stage holds once the copy has landed
    stage_pending = [None, None]  # tile whose copy is in flight
    reading_stage = [set(), set()]  # warps currently reading stage s
    reading_c = [set(), set()]
    writing_c = [set(), set()]
    barrier_count = [0] * n_tiles
    issued = set()

    def issue(t):
        s = t & 1
        if reading_stage[s]:
            raise Hazard('bulk copy of tile %d into stage %d while warps %s still read it' % (t, s, reading_stage[s]))
        stage_full[s] = None
        stage_pending[s] = t
        issued.add(t)

    issue(0)  # prologue: tile 0 only (the one-barrier form issues t + 1 inside the loop)
    for _ in range(max_steps):
        if all(t >= n_tiles for t in tile):
            return True
        # asynchronous copy engine
        for s in (0, 1):
            if stage_pending[s] is not None and rng.random() < 0.3:
                stage_full[s] = stage_pending[s]
                stage_pending[s] = None
        w = rng.randrange(n_warps)
        t = tile[w]
        if t >= n_tiles:
            continue
        s = t & 1
        if phase[w] == 0:
            if stage_full[s] == t:
                phase[w] = 1
                reading_stage[s].add(w)
                if reading_c[s]:
                    raise Hazard('warp %d writes C[%d] for tile %d while %s still read it' % (w, s, t, reading_c[s]))
                writing_c[s].add(w)
        elif phase[w] == 1:
            if rng.random() < 0.5:  # GEMM1 + transform done
                writing_c[s].discard(w)
                phase[w] = 2
                barrier_count[t] += 1
        elif phase[w] == 2:
            if barrier_count[t] == n_warps:  # barrier released
                if w == 0 and t + 1 < n_tiles and (t + 1) not in issued:
                    issue(t + 1)
                if w != 0 or t + 1 >= n_tiles or (t + 1) in issued:
                    phase[w] = 3
                    if writing_c[s]:
                        raise Hazard('warp %d reads C[%d] while it is being written' % (w, s))
                    reading_c[s].add(w)
        elif phase[w] == 3:
            if rng.random() < 0.5:  # GEMM2 done
                reading_c[s].discard(w)
                reading_stage[s].discard(w)
                tile[w] = t + 1
                phase[w] = 0
    return False


@pytest.mark.parametrize('seed', range(20))
def test_one_barrier_loop_has_no_reuse_hazard(seed):
    assert run(n_tiles=9, n_warps=8, seed=seed)


def test_model_detects_a_missing_barrier():
    """The same loop with the stage refilled two tiles ahead (the round-1 schedule) but WITHOUT its second barrier must
    trip the stage hazard: the model is able to see the bug it guards against."""

    def broken(seed):
        rng = random.Random(seed)
        n_warps, n_tiles = 8, 9
        tile = [0] * n_warps
        busy = [False] * n_warps
        reading = [set(), set()]
        for _ in range(20000):
            w = rng.randrange(n_warps)
            t = tile[w]
            if t >= n_tiles:
                continue
            s = t & 1
            if not busy[w]:
                busy[w] = True
                reading[s].add(w)
            elif rng.random() < 0.5:
                busy[w] = False
                reading[s].discard(w)
                tile[w] = t + 1
                if w == 0 and t + 2 < n_tiles and reading[s]:  # refill without waiting for the other warps
                    return True
        return False

    assert any(broken(seed) for seed in range(5))
