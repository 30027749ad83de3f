"""Multi-GPU plumbing for the decode path: codewords are independent, so a batch is split into contiguous shards,
one per rank (one process per GPU), with NO collective on the data path.  The only exchanges are the tiny counter /
gradient all-reduces below (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

from . import ops  # noqa: F401  (registers torch.ops.nldpc.*)


def shard_bounds(total: int, world: int, rank: int):
    """contiguous [begin, end) of `total` codewords owned by `rank` (sizes differ by at most one)"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor):
    """in-place SUM all-reduce when a process group exists (no-op single process)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_mean_grads_(params):
    """data-parallel training: ONE all-reduce of the flat weight-gradient vector, then divide by the world size
    (the loss is a mean over equal shards, LDPCDecoderLoss.py:108)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def count_errors_packed(hard: torch.Tensor, expected_bits_packed: torch.Tensor):
    """bit / frame error counts from packed hard decisions [B, nbytes] (uint8) against packed expected bits.
    Returns int64 tensor [bit_errors, frame_errors, bits, frames] on hard's device (all-reduce it across ranks)."""
    if hard.is_cuda:            # fused kernel (nldpc_count_errors_packed): one pass, no [B, nbytes] int64 temporaries
        c = torch.ops.nldpc.count_errors_packed(hard, hard.shape[1] * 8, expected_bits_packed)
        return torch.cat([c[:, 0], torch.tensor([hard.shape[0] * hard.shape[1] * 8, hard.shape[0]], dtype=torch.int64).to(hard.device)])
    x = torch.bitwise_xor(hard, expected_bits_packed)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int64, device=hard.device)
    per_cw = lut[x.long()].sum(dim=1)
    return torch.stack([per_cw.sum(), (per_cw > 0).sum(), torch.tensor(hard.shape[0] * hard.shape[1] * 8, device=hard.device),
                        torch.tensor(hard.shape[0], device=hard.device)]).to(torch.int64)
