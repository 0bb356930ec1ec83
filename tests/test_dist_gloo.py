"""World-size-2 test of the multi-GPU host logic on CPU (gloo): the
variable-length block-list exchange that precedes the album union query
(loudgain_b200.engine.gather_block_lists), checked end to end against the
oracle's album result on the union of both ranks' tracks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gated(z):
    """Integrated loudness of a block list, numpy."""
    gate = 10 ** ((-70 + 0.691) / 10)
    z = z[z >= gate]
    if not len(z):
        return -np.inf
    thr = 0.1 * z.mean()
    z = z[z >= thr]
    return 10 * np.log10(z.mean()) - 0.691


def _worker(rank, world, port, blocks, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from loudgain_b200.engine import gather_block_lists
        local = torch.from_numpy(blocks[rank])
        lists = gather_block_lists(dist, local, world)
        assert [len(x) for x in lists] == [len(b) for b in blocks]
        for got, want in zip(lists, blocks):
            np.testing.assert_array_equal(got.numpy(), want)
        out[rank] = _gated(np.concatenate([x.numpy() for x in lists]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_album_union_across_two_ranks(oracle):
    from oracle import blocks as oracle_blocks
    rng = np.random.default_rng(2)
    pcm = [(rng.standard_normal((44100 * (5 + 3 * i), 2)) * (800 * (i + 1))).astype(np.int16)
           for i in range(3)]
    states = []
    for p in pcm:
        st = oracle.init(2, 44100)
        st.add_frames(p, 4096)
        states.append(st)
    want = oracle.loudness_global_multiple(states)
    # rank 0 holds tracks 0 and 2, rank 1 holds track 1 (uneven list lengths)
    per_rank = [np.concatenate([oracle_blocks(oracle, states[0], 0), oracle_blocks(oracle, states[2], 0)]),
                oracle_blocks(oracle, states[1], 0)]
    for st in states:
        st.destroy()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), per_rank, out), nprocs=2, join=True)
    assert abs(out[0] - want) < 1e-9 and abs(out[1] - want) < 1e-9


@pytest.mark.timeout(120)
def test_empty_rank_contributes_nothing():
    per_rank = [np.array([1e-3, 2e-3, 5e-4]), np.zeros(0)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), per_rank, out), nprocs=2, join=True)
    assert out[0] == out[1] == _gated(per_rank[0])


def _segment_worker(rank, world, port, slots, rate, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from loudgain_b200.engine import gather_block_lists
        from tests.helpers import gate_slots
        lists = gather_block_lists(dist, torch.from_numpy(slots[rank]), world)
        peaks = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(peaks, op=dist.ReduceOp.MAX)
        out[rank] = gate_slots(np.concatenate([x.numpy() for x in lists]), rate) + (float(peaks[0]),)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_time_sharded_stream_across_two_ranks(oracle):
    """BASELINE config 4 in small: one stream, two ranks, each sweeping its time
    segment (one second of lead-in instead of a filter-state exchange, host
    emulation of the device math), the 100 ms slot energies all-gathered in rank
    = time order, blocks and gates formed over the whole list on every rank."""
    from loudgain_b200 import synth
    from loudgain_b200.engine import segment_plan
    from tests.helpers import emu_measure, oracle_measure
    rate = 48000
    spec = synth.config1_spec(31.7)
    spec.rate = rate
    pcm = synth.programme_s16(spec).numpy()
    want = oracle_measure(oracle, [(pcm, rate)])["tracks"][0]
    s100 = (rate + 5) // 10
    plan = segment_plan(len(pcm), rate, 2)
    per_rank = []
    for a, lead, e in plan:
        r = emu_measure([(pcm[a:e], rate)], lead_in=[lead])
        per_rank.append(np.ascontiguousarray(r["slots"][0][lead // s100:]))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_segment_worker, args=(2, _free_port(), per_rank, rate, out), nprocs=2, join=True)
    for rank in range(2):
        loud, rng, pk = out[rank]
        assert abs(loud - want["loudness"]) < 2e-4 and abs(rng - want["range"]) < 2e-4
        assert pk == 2.0
    assert out[0] == out[1]
