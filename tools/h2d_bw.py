#!/usr/bin/env python
"""Raw pinned host -> device copy bandwidth of this box (the ceiling of bench.py's e2e number).

    python tools/h2d_bw.py                                                       # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bw.py
        # N ranks: every rank ALONE (the others idle), then ALL ranks at once -> what the platform gives N concurrent
        # host->device streams (shared root complexes / host memory / hypervisor), independent of any decode work
"""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
n = 65536 * 832
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device=dev)


def measure(reps=10):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return n * 4 * reps / (time.perf_counter() - t0) / 1e9


if world == 1:
    gbs = measure()
    print(f"H2D {n * 4 / 1e6:.0f} MB pinned: {gbs:.1f} GB/s = {gbs * 1e9 / 3328 / 1e6:.2f} M BG2 codewords/s (fp32 LLRs, 3328 B per codeword)")
else:
    alone = torch.zeros(world, device=dev)
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone[r] = measure()
        dist.barrier()
    dist.all_reduce(alone)
    dist.barrier()
    together = torch.zeros(world, device=dev)
    together[rank] = measure()
    dist.all_reduce(together)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "pinned_MB_per_copy": n * 4 / 1e6,
                          "h2d_GBps_each_rank_alone": [round(float(v), 1) for v in alone],
                          "h2d_GBps_all_ranks_concurrently": [round(float(v), 1) for v in together],
                          "aggregate_concurrent_GBps": round(float(together.sum()), 1),
                          "bg2_codewords_per_s_ceiling_fp32_llrs": round(float(together.sum()) * 1e9 / 3328),
                          "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."], "cpus": len(os.sched_getaffinity(0))}))
    dist.destroy_process_group()
