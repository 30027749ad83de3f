#!/usr/bin/env python
"""BASELINE config 5: train_BoostedNeuralLDPCDecoder step (forward + multi-iteration BCE + backward + clip + Adam + clamp),
BG2 z16, QMS q=5, cn=3 / vn=3, T=20, data-parallel with ONE NCCL all-reduce of the weight gradients per step.

    python tools/bench_train.py --batch 4096 --steps 20
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py --batch 4096
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType  # noqa: E402
from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig  # noqa: E402
from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator, FusedTrainer, train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096, help="codewords per GPU per step (the reference uses 20)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--mode", default="torch", choices=["torch", "fused", "graph"],
                    help="torch: clip_grad_norm_ / torch.optim.Adam / clamp as the reference calls them; fused: FusedTrainer (flat "
                         "vector, one optimiser launch); graph: the same, whole step replayed from a CUDA graph")
    ap.add_argument("--fixed-batch", action="store_true", help="reuse one batch (leaves the torch-op batch generator out of the step)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bg, Z = load_basegraph("nr_bg2_set0")
    graph = TannerGraph(bg, Z)
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    T, B = args.iters, args.batch
    model = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(3, 0, 3),
                                     decoding_type=DecoderType.QMS, decoder_qms_qbit=5).to(dev)
    model.store_llr = "none"
    crit = LDPCDecoderLoss(LossType.BCE, etha=1.0)
    opt = torch.optim.Adam(model.get_trainable_parameters(), lr=1e-3)
    gen = DeviceBatchGenerator(graph, [2, 2.5, 3.0, 3.5, 4.0], dev, seed=2042 + rank, qms_qbit=5)

    trainer = FusedTrainer(model, crit, T, lr=1e-3, graph=(args.mode == "graph")) if args.mode != "torch" else None
    fixed = gen(B) if args.fixed_batch else None

    def step():
        x, y = fixed if fixed is not None else gen(B)
        if trainer is not None:
            return trainer.step(x, y)
        return train_step(model, crit, opt, x, y, T)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(args.steps):
        losses.append(step())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "train_step_codewords_per_s", "value": B * world * args.steps / (ms * 1e-3), "unit": "codewords/s",
                          "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps, "batch_per_gpu": B, "iterations": T, "mode": args.mode,
                          "fixed_batch": bool(args.fixed_batch),
                          "config": "BoostedNeuralLDPCDecoder BG2 z16 QMS5 cn3/vn3, BCE etha=1, clip 1.0, Adam 1e-3, clamp [0,2]",
                          "first_loss": float(losses[0].detach()), "last_loss": float(losses[-1].detach())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
