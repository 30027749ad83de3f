#!/usr/bin/env python
"""Time the STOCK reference (unmodified PyTorch code, CPU) on this machine's host cores — bench.py's `cpu_baseline_stock`.

Runs in its own process (bench.py spawns it) because the reference's package names (`neural_ldpc_decoder`, ...) are the
names the drop-in mirrors can register themselves under.  Imports the reference from baseline/_ref (placed by
tools/install_reference.sh; travels to the GPU box) or, in the build container, from /root/reference/src.  Nothing of this
repo's product code or oracle is imported: only the base-graph JSON files are read.

Workloads (SURVEY.md §8(d) "Reference CPU timing beside it"): BASELINE configs[0] verbatim — NeuralLDPCDecoder WiMAX z=24,
batch 1024, 10 iterations — and the headline code at the largest batch the dense formulation handles comfortably: BG2 z=16,
batch 256, 10 iterations (2.48 MB of temporaries per codeword).  torch.no_grad(), all host threads, one small warm-up call,
then one timed call each (`--repeats` for more; the best is reported).  Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = os.path.join(ROOT, "neural_ldpc_decoder_torch_b200", "resources")


def load_bg(name):
    with open(os.path.join(RES, name + ".json")) as f:
        doc = json.load(f)
    bg = -np.ones((doc["M"], doc["N"]), dtype=np.int64)
    for i, row in enumerate(doc["rows"]):
        for j, s in row:
            bg[i, j] = s
    return bg, int(doc["Z_default"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--bg2-batch", type=int, default=256)
    ap.add_argument("--wimax-batch", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    src = None
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "neural_ldpc_decoder")):
            src = cand
            break
    if src is None:
        print(json.dumps({"unavailable": "no reference under baseline/_ref or /root/reference/src (run tools/install_reference.sh)"}))
        return
    sys.dont_write_bytecode = True
    sys.path.insert(0, src)
    import torch
    import neural_ldpc_decoder as nref

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    out = {"source": os.path.relpath(src, ROOT) if src.startswith(ROOT) else src, "torch": torch.__version__, "cpu_count": os.cpu_count(),
           "affinity": cores, "torch_threads": torch.get_num_threads(), "runs": []}
    T = args.iters
    for name, label, sigma, B in (("wimax_n576_r34", "NeuralLDPCDecoder WiMAX N=576 R=3/4 z=24 (BASELINE configs[0] verbatim)", 0.62095, args.wimax_batch),
                                  ("nr_bg2_set0", "NeuralLDPCDecoder 5G NR BG2 z=16 (headline code)", 1.2559, args.bg2_batch)):
        bg, Z = load_bg(name)
        M, N = bg.shape
        cm = nref.ConnectingMatrixTorch(nref.ConnectingMatrix(Z=Z, basegraph=bg))
        model = nref.NeuralLDPCDecoder(T, B, cm)
        rs = np.random.RandomState(7)
        E = int((bg != -1).sum())
        with torch.no_grad():
            for t in range(T):      # the bench's trained-like weights (timing does not depend on them; kept for like-for-like work)
                model.weights_var[t].copy_(torch.from_numpy(rs.uniform(0.3, 1.3, E).astype(np.float32)))
                model.biases_var[t].copy_(torch.from_numpy((0.2 * rs.normal(size=E)).astype(np.float32)))
            xa = torch.from_numpy((2.0 * (sigma * rs.normal(0, 1, (B, N, Z)) - 1.0) / sigma ** 2).astype(np.float32))
            model(xa[:8])          # warm-up (thread pool, allocator)
            best = None
            for _ in range(max(1, args.repeats)):
                t0 = time.perf_counter()
                outs = model(xa)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            assert len(outs) == T and tuple(outs[-1].shape) == (B, N * Z)
        out["runs"].append({"workload": f"{label}, batch {B}, {T} iterations, forward() on CPU, torch {torch.__version__}, {cores} threads",
                            "value": B / best, "unit": "codewords/s", "gbit_per_s": B / best * N * Z / 1e9, "seconds": best, "batch": B})
        del model, cm
    print(json.dumps(out))


if __name__ == "__main__":
    main()
