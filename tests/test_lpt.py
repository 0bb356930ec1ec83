"""Host logic of the library scan over several GPUs: the track-to-rank
assignment (lgb_lpt_assign; the reference lets the OS balance one process per
file / album, bin/rgbpm2:150-175).  No GPU involved."""
import random

from loudgain_b200 import engine


def _loads(costs, ranks, world):
    loads = [0] * world
    for c, r in zip(costs, ranks):
        loads[r] += c
    return loads


def test_lpt_balances_a_library():
    rng = random.Random(7)
    # cfg5-like: 10 000 stereo tracks of 2-8 minutes at 44.1 kHz
    costs = [2 * int(rng.uniform(120, 480) * 44100) for _ in range(10000)]
    for world in (1, 2, 4, 8):
        ranks, worst = engine.lpt_assign(costs, world)
        loads = _loads(costs, ranks, world)
        assert max(loads) == worst and sum(loads) == sum(costs)
        # LPT bound: within one smallest item of the mean once items are plentiful
        assert max(loads) - min(loads) <= min(costs)
        assert max(loads) <= sum(costs) / world * 1.0005


def test_lpt_is_deterministic_and_handles_edges():
    costs = [5, 5, 5, 9, 1, 9, 0]
    a, worst = engine.lpt_assign(costs, 3)
    b, _ = engine.lpt_assign(costs, 3)
    assert a == b                                  # every rank computes the same plan
    assert worst == max(_loads(costs, a, 3)) == 14
    assert a[3] == 0 and a[5] == 1                 # equal costs keep their order, ties go to the lower rank
    assert engine.lpt_assign([], 4) == ([], 0)
    ranks, worst = engine.lpt_assign([7, 3], 8)    # more ranks than items
    assert sorted(ranks) == [0, 1] and worst == 7


def test_bench_union_gating_is_the_oracles(oracle):
    """bench.py checks the multi-GPU album against a numpy gating of the union of all
    ranks' block lists (merged_album_check); that restatement must be the oracle's rule."""
    import numpy as np

    import bench
    from loudgain_b200 import synth
    from oracle import blocks as oracle_blocks

    specs = synth.config2_specs(ntracks=3, scale=0.1)
    sts, z, st = [], [], []
    for s in specs:
        state = oracle.init(2, s.rate)
        state.add_frames(synth.programme_s16(s).numpy(), 4096)
        sts.append(state)
        z.append(oracle_blocks(oracle, state, 0))
        st.append(oracle_blocks(oracle, state, 1))
    want_l, want_r = oracle.loudness_global_multiple(sts), oracle.loudness_range_multiple(sts)
    got_l, got_r = bench._numpy_gating(np.concatenate(z), np.concatenate(st))
    for state in sts:
        state.destroy()
    assert abs(got_l - want_l) < 1e-9 and abs(got_r - want_r) < 1e-9
