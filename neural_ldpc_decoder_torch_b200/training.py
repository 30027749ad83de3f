"""Training-step plumbing around the CUDA decode/backward kernels — the B200 equivalent of the loop body of
/root/reference/train/train_BoostedNeuralLDPCDecoder.py:260-294:

    batch -> model(x, target_iter=range(T)) -> LDPCDecoderLoss -> backward -> [NCCL all-reduce of the weight gradients]
          -> clip_grad_norm_(1.0) -> Adam -> model._apply_constraints()

plus an on-device replacement of the reference's O(B^2) numpy batch generator (boosted AWGNPassedDatagen.py:136-193):
random information words, systematic encoding with the generator matrix, BPSK + AWGN at mixed SNRs, LLR, QMS
pre-quantisation — all as a handful of torch ops on the GPU (plumbing, not the hot path)."""
import math

import numpy as np
import torch

from .sharding import allreduce_mean_grads_


class DeviceBatchGenerator:
    """Seeded mix-SNR batch generator on `device` (same distribution as the reference's `mix_snr` generator, its own
    random stream).  Returns x [B, N, Z] float32 (LLR, bit 1 <-> positive) and y [B, N*Z] float32 (codeword bits)."""

    def __init__(self, graph, snr_db, device, seed=2042, gen_matrix=None, all_zero=False, qms_qbit=None, rate_denominator_minus=2):
        self.g = graph
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        K = graph.N - graph.M
        rate = 1.0 * K / (graph.N - rate_denominator_minus)      # the reference's K / (N - 2)
        snr_lin = 10.0 ** (np.asarray(snr_db, dtype=np.float64) / 10.0)
        self.sigmas = torch.tensor(np.sqrt(1.0 / (2.0 * snr_lin * rate)), dtype=torch.float32, device=self.device)
        self.all_zero = all_zero
        self.qms_qbit = qms_qbit
        self.G = None
        if not all_zero:
            G = graph.systematic_generator() if gen_matrix is None else np.asarray(gen_matrix)
            self.G = torch.tensor(G, dtype=torch.float32, device=self.device)      # [K*Z, N*Z] 0/1

    @torch.no_grad()
    def __call__(self, B):
        g, dev = self.g, self.device
        NZ = g.N * g.Z
        if self.all_zero:
            y = torch.zeros((B, NZ), dtype=torch.float32, device=dev)
        else:
            info = torch.randint(0, 2, (B, self.G.shape[0]), generator=self.gen, device=dev).float()
            y = torch.remainder(info @ self.G, 2.0)          # exact: sums <= K*Z < 2^24
        sigma = self.sigmas[torch.arange(B, device=dev) % self.sigmas.numel()].unsqueeze(1)
        noise = torch.randn((B, NZ), generator=self.gen, device=dev)
        received = noise * sigma + (2.0 * y - 1.0)           # bit 0 -> -1, bit 1 -> +1
        x = 2.0 * received / (sigma * sigma)
        if self.qms_qbit == 5:
            x = torch.clamp(torch.round(x * 2.0) / 2.0, -7.5, 7.5)
        elif self.qms_qbit is not None:
            from .boosted_neural_ldpc_decoder.Functions import Functions
            x = Functions.cal_msa_q_torch(x, self.qms_qbit)
        return x.reshape(B, g.N, g.Z).contiguous(), y


def train_step(model, criterion, optimizer, x, y, n_iters, max_grad_norm=1.0):
    """one optimisation step, train/train_BoostedNeuralLDPCDecoder.py:274-294; data-parallel when a process group exists."""
    model.train()
    optimizer.zero_grad()
    outputs = model(x, target_iter=list(range(n_iters)))
    loss = criterion(outputs, y, coeff_param=list(range(len(outputs))))
    loss.backward()
    allreduce_mean_grads_(list(model.parameters()))
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm)
    optimizer.step()
    model._apply_constraints()
    return loss


class FusedTrainer:
    """The loop body of train/train_BoostedNeuralLDPCDecoder.py:274-294 with the optimiser tail in ONE launch and, optionally,
    the whole step replayed from a CUDA graph.

    * The trainable parameters (model.get_trainable_parameters() / model.parameters()) are re-pointed at slices of one flat
      fp32 vector, their .grad at slices of a second one: the data-parallel exchange is ONE all-reduce (SUM) of that vector
      with no gather / scatter copies, and clip_grad_norm_ + Adam + _apply_constraints is one nldpc_clip_adam_clamp launch
      (1 / world_size folded in).  Parameter names, shapes and state_dict() are unchanged.
    * graph=True captures zero-grad -> forward -> fused BCE -> backward -> [all-reduce] -> optimiser once and replays it per
      step (inputs are copied into static buffers): for the reference's own batch size (20) the step is host-launch bound
      (~2.7 ms with torch ops, 1.9 ms eager here), the replay (1.1 ms) removes that.  While capturing, the
      specialised kernels take their weights from a fixed constant-arena range, ConstArena::acquire_captured).  Works under
      torch.distributed too (the NCCL all-reduce of the flat gradient is captured with the step; 1.19 ms per B = 20 step on 2
      GPUs) — call close() before dist.destroy_process_group(): tearing the communicator down while a graph that contains its
      kernels is alive hangs (tools/repro_nccl_graph.py)."""

    def __init__(self, model, criterion, n_iters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0, clamp=None,
                 params=None, graph=False):
        self.model, self.criterion, self.n_iters = model, criterion, int(n_iters)
        self.lr, self.betas, self.eps, self.max_grad_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(max_grad_norm)
        if clamp is None:
            rng = getattr(model, "allowed_weight_range", None)
            clamp = (float(rng.start), float(rng.end)) if rng is not None else (-float("inf"), float("inf"))
        self.clamp = clamp
        if params is None:
            params = model.get_trainable_parameters() if hasattr(model, "get_trainable_parameters") else model.parameters()
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        # parameters the optimiser does not update (fixed_iterative_nodes_init_weight > 0) still receive gradients, and the
        # reference clips over model.parameters() (train/...py:291): their gradients live behind the trainable ones in the flat
        # gradient vector, enter the norm and are scaled, but no weight moves
        own = {id(p) for p in self.params}
        self.frozen = [p for p in model.parameters() if p.requires_grad and id(p) not in own]
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params + self.frozen):
            from ._lib import NldpcError
            raise NldpcError("FusedTrainer needs fp32 parameters on one CUDA device (there is no CPU fallback)")
        n = sum(p.numel() for p in self.params)
        n_all = n + sum(p.numel() for p in self.frozen)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n_all, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)          # [step count, last gradient norm]
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)   # read by the kernel at run time (set_lr)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                off += k
            for p in self.frozen:
                k = p.numel()
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                off += k
        self.device = dev
        self.fused_loss = True       # use model.fused_bce_loss (one launch for forward + BCE + dL/dout) where it is covered
        self._graph = None
        self._want_graph = bool(graph)
        self._static = None

    # -- schedule / checkpoint ------------------------------------------------------------------------------------
    def set_lr(self, lr):
        """learning-rate schedule (train/...py:262-266 sets param_group['lr'] per epoch): the kernel reads the rate from a device
        scalar, so a captured step follows it too"""
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    def state_dict(self):
        """optimiser state for CheckPointUtil.save(optimizer=trainer): Adam moments per parameter, step count, lr"""
        return {"step": int(self.state[0].item()), "lr": self.lr, "betas": self.betas, "eps": self.eps,
                "exp_avg": self.exp_avg.detach().cpu().clone(), "exp_avg_sq": self.exp_avg_sq.detach().cpu().clone(),
                "param_numels": [p.numel() for p in self.params]}

    def load_state_dict(self, sd):
        if list(sd["param_numels"]) != [p.numel() for p in self.params]:
            raise ValueError("optimizer state belongs to a different parameter set")
        with torch.no_grad():
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.state[0] = float(sd["step"])
        self.betas, self.eps = (float(sd["betas"][0]), float(sd["betas"][1])), float(sd["eps"])
        self.set_lr(sd["lr"])

    # -- pieces ---------------------------------------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def _step_body(self, x, y):
        from . import ops
        self.flat_grad.zero_()                                            # grads accumulate in place into the flat vector
        loss = None
        if self.fused_loss and hasattr(self.model, "fused_bce_loss") and self._criterion_is_fused_bce() \
                and self.n_iters == int(self.model.iter_node_counts):
            # forward + multi-iteration BCE + dL/dout in ONE launch (no [T, B, N*Z] tensor crosses HBM twice); None = not covered
            loss = self.model.fused_bce_loss(x, y, etha=self.criterion.etha, coeff_param=list(range(self.n_iters)))
        if loss is None:
            outputs = self.model(x, target_iter=list(range(self.n_iters))) if self._takes_target_iter() else self.model(x)
            loss = self.criterion(outputs, y, coeff_param=list(range(len(outputs))))
        loss.backward()
        world = self._world()
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)        # the one exchange of the step (NCCL over NVLink)
        ops.clip_adam_clamp_(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.state, 1.0 / world, self.max_grad_norm,
                             self.lr, self.betas, self.eps, self.clamp, lr_dev=self.lr_dev)
        return loss.detach()

    def _criterion_is_fused_bce(self):
        from .boosted_neural_ldpc_decoder.struct.LossType import LossType
        return getattr(self.criterion, "loss_type", None) == LossType.BCE and bool(getattr(self.criterion, "fused", False))

    def _takes_target_iter(self):
        return hasattr(self.model, "fetch_param")                        # the Boosted decoder; the Neural forward takes xa only

    def _check_grad_views(self):
        off = 0
        for p in self.params + self.frozen:
            k = p.numel()
            is_param = off < self.flat.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off or (
                    is_param and p.data_ptr() != self.flat.data_ptr() + 4 * off):
                raise RuntimeError("a parameter or its .grad no longer aliases the trainer's flat vectors (zero_grad(set_to_none=True) "
                                   "or .to() was called on the model): build a new FusedTrainer")
            off += k

    # -- public ---------------------------------------------------------------------------------------------------
    def step(self, x, y):
        """one optimisation step on (x [B,N,Z], y [B,N*Z]); returns the (detached) loss tensor, no host synchronisation"""
        self.model.train()
        if not self._want_graph:
            self._check_grad_views()
            return self._step_body(x, y)
        if self._graph is None:
            self._check_grad_views()
            self._static = (torch.empty_like(x), torch.empty_like(y))
            self._static[0].copy_(x)
            self._static[1].copy_(y)
            saved = [t.clone() for t in (self.flat, self.exp_avg, self.exp_avg_sq, self.state)]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                                 # warm-up on a side stream (allocator, caches, NCCL)
                for _ in range(2):
                    self._step_body(*self._static)
            torch.cuda.current_stream(self.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._static_loss = self._step_body(*self._static)
            self._graph = graph
            for t, s0 in zip((self.flat, self.exp_avg, self.exp_avg_sq, self.state), saved):
                t.copy_(s0)                                               # the warm-up steps must not count: restore, then replay once
            self._graph.replay()
            return self._static_loss
        if x.shape != self._static[0].shape or y.shape != self._static[1].shape:
            raise ValueError("a graphed FusedTrainer is bound to the batch shape it was captured with")
        self._static[0].copy_(x)
        self._static[1].copy_(y)
        self._graph.replay()
        return self._static_loss

    def close(self):
        """release the captured step.  Under torch.distributed call this BEFORE dist.destroy_process_group(): tearing the NCCL
        communicator down while a CUDA graph that contains its kernels is alive hangs (tools/repro_nccl_graph.py)."""
        self._graph = None
        self._static = None
        self._static_loss = None

    @property
    def steps_done(self):
        return int(self.state[0].item())

    @property
    def last_grad_norm(self):
        return float(self.state[1].item())


def wilson_interval(k, n, z=1.959963984540054):
    """95 % Wilson score interval of a binomial proportion (BER/FER curves are compared through it)"""
    if n == 0:
        return 0.0, 1.0
    p = k / n
    den = 1.0 + z * z / n
    c = (p + z * z / (2 * n)) / den
    h = z * math.sqrt(p * (1 - p) / n + z * z / (4.0 * n * n)) / den
    return max(0.0, c - h), min(1.0, c + h)
