#!/usr/bin/env python
"""Minimal repro: a NCCL all-reduce inside a captured CUDA graph (what FusedTrainer(graph=True) would need under
torch.distributed).  Run under torchrun with 2 ranks, each variant under its own `timeout -s KILL`:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/repro_nccl_graph.py plain
    ... tools/repro_nccl_graph.py trainer        # the FusedTrainer step itself, graph=True

prints one JSON line per rank-0 run; a hang shows as the timeout killing the job.

Measured (torch 2.11.0+cu128, NCCL 2.28.9, 2 x B200): the captured all-reduce and the captured trainer step both WORK (ranks' weights
stay bit-identical, 1.19 ms per B = 20 step); what hangs is dist.destroy_process_group() while a CUDA graph that contains NCCL kernels
is still alive (KEEP_GRAPH=1 reproduces it).  Release the graph first: FusedTrainer.close()."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
variant = sys.argv[1] if len(sys.argv) > 1 else "plain"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"variant": variant, "torch": torch.__version__, "nccl": ".".join(map(str, torch.cuda.nccl.version())), "world": world}
t0 = time.time()
if variant == "plain":
    x = torch.full((40,), float(rank + 1), device=dev)
    dist.all_reduce(x)                                    # communicator set up eagerly
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            dist.all_reduce(x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    x.fill_(float(rank + 1))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        dist.all_reduce(x)
        x.mul_(0.5)
    for _ in range(3):
        x.fill_(float(rank + 1))
        g.replay()
    torch.cuda.synchronize()
    out["result"] = float(x[0])                           # 0.5 * (1 + 2) = 1.5 with 2 ranks
    if os.environ.get("KEEP_GRAPH") != "1":
        del g                                             # a live graph that holds NCCL kernels makes destroy_process_group() hang
else:
    from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
    from neural_ldpc_decoder_torch_b200 import boosted_neural_ldpc_decoder as bn
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator, FusedTrainer
    bg, Z = load_basegraph("nr_bg2_set0")
    graph = TannerGraph(bg, Z)
    cm = bn.ConnectingMatrixTorch(bn.ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    B, T = 20, 20
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 3), decoding_type=DecoderType.QMS).to(dev)
    m.store_llr = "none"
    x, y = DeviceBatchGenerator(graph, [2.0, 3.0], dev, seed=5 + rank, qms_qbit=5)(B)
    tr = FusedTrainer(m, LDPCDecoderLoss(LossType.BCE, etha=1.0), T, graph=True)
    losses = [float(tr.step(x, y)) for _ in range(5)]
    torch.cuda.synchronize()
    w = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    out["losses"] = losses
    out["weights_identical_across_ranks"] = bool(all(torch.equal(ws[0], v) for v in ws))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        tr.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    out["ms_per_step"] = e0.elapsed_time(e1) / 50
    if os.environ.get("KEEP_GRAPH") != "1":
        tr.close()
out["seconds"] = round(time.time() - t0, 2)
if rank == 0:
    print(json.dumps(out), flush=True)
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print(json.dumps({"variant": variant, "clean_exit": True, "kept_graph": os.environ.get("KEEP_GRAPH") == "1"}), flush=True)
