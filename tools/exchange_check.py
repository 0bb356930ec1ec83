#!/usr/bin/env python3
"""Multi-GPU check of the album exchange (engine.AlbumExchange), one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29611 tools/exchange_check.py

Albums are dealt out over the ranks track by track (lgb_lpt_assign), every rank
measures its share with the exchange attached, and rank 0 also measures ALL tracks
alone: the exchanged album results must equal the single-GPU ones (loudness to
1e-10 LU: the partial sums meet in a different order; range and the gate counts
exactly), on every rank, on every repeat.  Prints one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loudgain_b200 import engine, synth  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nalbums, per_album = 5, 7
    specs = []
    for a in range(nalbums):
        ss = synth.config2_specs(per_album, scale=0.06 + 0.01 * a)
        for s in ss:
            s.seed += 100 * a
        specs += ss
    albums = [i // per_album for i in range(len(specs))]
    albums[-1] = engine.NO_ALBUM                      # a track outside every album
    costs = [2 * int(s.seconds * s.rate) for s in specs]
    owner, _ = engine.lpt_assign(costs, world)
    mine = [i for i in range(len(specs)) if owner[i] == rank]
    tracks = [(synth.programme_s16(specs[i], device=dev), specs[i].rate) for i in mine]
    b = engine.Batch(tracks, [albums[i] for i in mine], nalbums=nalbums)
    x = engine.AlbumExchange(b, dist if world > 1 else None, world, rank)
    runs = []
    for _ in range(5):
        b.run()
        _, ares = b.fetch()
        runs.append([(m.loudness, m.range, m.n_abs, m.n_rel, m.n_shortterm) for m in ares])
    ok = all(r == runs[0] for r in runs)
    worst = 0.0
    if world > 1:
        # every rank got the same bits
        t = torch.tensor([[v[0], v[1]] for v in runs[0]], dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ok = ok and all(torch.equal(allt[0], y) for y in allt)
    if rank == 0:
        full = [(synth.programme_s16(s, device=dev), s.rate) for s in specs]
        _, want = engine.measure(full, albums)
        for a in range(nalbums):
            g, w = runs[0][a], want[a]
            d = abs(g[0] - w.loudness)
            worst = max(worst, d)
            ok = ok and d <= 1e-10 and g[1] == w.range and g[2:] == (w.n_abs, w.n_rel, w.n_shortterm)
        print(json.dumps({"exchange_check": "ok" if ok else "FAILED", "world": world,
                          "max_loudness_diff_lu": worst,
                          "albums": [[v[0], v[1]] for v in runs[0]]}))
    b.close()
    x.close()
    if world > 1:
        ok_t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        ok = bool(ok_t.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
