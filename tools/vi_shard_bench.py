#!/usr/bin/env python3
"""Sharded soft-VI sweeps of ONE grid on N GPUs (torchrun): tunes the tile shape across the ranks, then reports us per sweep,
the halo waits of the edge CTAs, and the bits against the single-GPU sweeps.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/vi_shard_bench.py [grid4096] [f32fast]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import piql_jl_b200 as piql
    from piql_jl_b200 import distributed as pd
    name = sys.argv[1] if len(sys.argv) > 1 else "grid4096"
    prec = sys.argv[2] if len(sys.argv) > 2 else "f32fast"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % lr))
    walls, goal, params = bench.make_world(name)
    eng = piql.Engine(lr)
    eng.load_grid(walls, goal, **params)
    stream = torch.cuda.ExternalStream(eng.stream, device="cuda:%d" % lr)
    env, results = pd.tune_sharded(eng, "nu", rank, world, precision=prec)
    nsw = 400
    eng.vi_prepare("nu", precision=prec)
    blob, cols = eng.vi_shard_begin(rank, world)
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    eng.vi_shard_connect(blobs)
    dist.barrier()
    with torch.cuda.stream(stream):
        eng.vi_sweeps(40)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        eng.vi_shard_wait_stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        eng.vi_sweeps(nsw)
        b.record(stream)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda:%d" % lr)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wait_us, waits = eng.vi_shard_wait_stats()
    plan = eng.vi_tile_plan("nu", prec)
    w = torch.tensor([wait_us, float(waits)], dtype=torch.float64, device="cuda:%d" % lr)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    dist.barrier()
    eng.vi_shard_end()
    if rank == 0:
        us = float(t[0]) * 1e3 / nsw
        print(json.dumps({"grid": name, "precision": prec, "gpus": world, "us_per_sweep": us, "pinned": env, "plan": plan,
                          "mean_halo_wait_us_per_edge_cta_round": float(w[0]) / max(float(w[1]), 1.0),
                          "rounds": nsw // plan["sweeps_per_round"],
                          "tuning_us_per_sweep": {json.dumps(dict(k)): round(v, 2) for k, v in sorted(results.items(), key=lambda kv: kv[1])}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
