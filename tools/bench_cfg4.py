#!/usr/bin/env python3
"""BASELINE config 4: ONE long 48 kHz stereo S16 stream, time-sharded across the
ranks of one box (python tools/bench_cfg4.py, or torchrun --nproc-per-node N).

The stream is the concatenation of 60 s programme pieces (piece i has seed
17704 + i), so every rank can synthesise exactly the part it owns -- its time
segment plus one second of lead-in (loudgain_b200.engine.segment_plan) -- on
its own GPU.  A step = sweep of the rank's segment, all-gather of the 100 ms
slot energies and peaks over NCCL, blocks + gating + range over the whole slot
list on every rank (engine.StreamShard).  With --check the result
is compared with rank 0 measuring the whole stream alone (use a --seconds that
fits one GPU).  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PIECE_S = 60.0
RATE = 48000


def stream_part(first: int, end: int, total: int, device):
    """Frames [first, end) of the stream, int16 [frames, 2] on `device`."""
    import torch
    from loudgain_b200 import synth
    piece = int(PIECE_S * RATE)
    out = []
    for i in range(first // piece, (end + piece - 1) // piece):
        n = min(piece, total - i * piece)
        spec = synth.TrackSpec(seed=17704 + i, rate=RATE, channels=2, seconds=n / RATE)
        p = synth.programme_s16(spec, device=device)[:n]
        lo, hi = max(first - i * piece, 0), min(end - i * piece, n)
        out.append(p[lo:hi])
    return torch.cat(out).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=36000.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from loudgain_b200 import build, engine
    build()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    total = int(args.seconds * RATE)
    first, lead, end = engine.segment_plan(total, RATE, world)[rank]
    seg = [(stream_part(first, end, total, dev), lead)]

    shard = engine.StreamShard(seg, RATE, dist if world > 1 else None, world)

    def step():
        shard.run()
        return shard.fetch()

    for _ in range(args.warmup):
        m = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    line = {"workload": f"cfg4: one {args.seconds:.0f} s 48 kHz stereo S16 stream, time-sharded",
            "n_gpus": world, "samples": 2 * total, "steps": args.steps,
            "ms_per_step": float(ms.item()) / args.steps,
            "value": 2 * total * args.steps / (float(ms.item()) * 1e-3) / 1e9, "unit": "Gsamples/s",
            "loudness": m.loudness, "range": m.range, "true_peak": [float(x) for x in m.true_peak],
            "sharding": "by time, 1 s lead-in per segment, slot energies all-gathered over NCCL"}
    if args.check and rank == 0:
        whole, _ = engine.measure([(stream_part(0, total, total, dev), RATE)])
        line["check"] = {"loudness_diff": abs(whole[0].loudness - m.loudness),
                         "range_diff": abs(whole[0].range - m.range),
                         "true_peak_equal": bool((whole[0].true_peak == m.true_peak).all()),
                         "sample_peak_equal": bool((whole[0].sample_peak == m.sample_peak).all())}
        # the whole stream on one GPU is planned with longer chunks than a segment,
        # so the FP32 filter rounds differently: equal to ~1e-5 LU, not bit for bit
        assert line["check"]["loudness_diff"] < 1e-4 and line["check"]["range_diff"] < 1e-4
        assert line["check"]["true_peak_equal"] and line["check"]["sample_peak_equal"]
    shard.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
