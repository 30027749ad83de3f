"""GPU: the fused tail of the training step (nldpc_clip_adam_clamp = clip_grad_norm_ + torch.optim.Adam.step +
_apply_constraints, train/train_BoostedNeuralLDPCDecoder.py:291-294) against the torch ops the reference calls, and the
FusedTrainer (eager and CUDA-graph replay) against training.train_step.  fp32 tolerance 2e-6 absolute on weights in [0, 2]
per step (different but equivalent operation order in the norm and the Adam update)."""
import copy

import numpy as np
import pytest
import torch

from boosted_util import build_module
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,max_norm", [(1, 1.0), (40, 1.0), (3940, 1.0), (1000, 0.0), (257, 1e-3)])
def test_clip_adam_clamp_matches_torch_ops(n, max_norm):
    from neural_ldpc_decoder_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    p0 = torch.rand(n, generator=g, device="cuda") * 2.0
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p, m, v, st = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), torch.zeros(2, device="cuda")
    world = 4
    for step in range(25):
        grad = torch.randn(n, generator=g, device="cuda") * (10.0 if step % 3 == 0 else 0.01)       # clipped and unclipped steps
        ref_p.grad = (grad / world).clone()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_([ref_p], max_norm=max_norm)
        else:
            total = ref_p.grad.norm()
        clipped = ref_p.grad.clone()
        opt.step()
        ref_p.data.clamp_(0.0, 2.0)
        gbuf = grad.clone()
        ops.clip_adam_clamp_(p, gbuf, m, v, st, grad_scale=1.0 / world, max_norm=max_norm, clamp=(0.0, 2.0))
        assert float(st[0]) == step + 1
        assert abs(float(st[1]) - float(total)) <= 1e-5 * float(total)
        assert torch.allclose(gbuf, clipped, rtol=1e-5, atol=1e-12)
        assert float((p - ref_p.data).abs().max()) < 2e-6 * (step + 1), (step, float((p - ref_p.data).abs().max()))
    assert float(p.min()) >= 0.0 and float(p.max()) <= 2.0


def _fresh(tag):
    d = load_golden(f"train_boosted_{tag}")
    return d, build_module(d, device="cuda")


@pytest.mark.parametrize("tag", ["cn2vn3_qms", "cn1vn2_ms"])
@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_follows_reference_training_loop(tag, graph):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from neural_ldpc_decoder_torch_b200.training import FusedTrainer, train_step
    d, m_ref = _fresh(tag)
    _, m_new = _fresh(tag)
    T = int(d["T"])
    crit = LDPCDecoderLoss(LossType.BCE, etha=float(d["etha"]))
    opt = torch.optim.Adam(m_ref.get_trainable_parameters(), lr=1e-3)
    tr = FusedTrainer(m_new, crit, T, lr=1e-3, graph=graph)
    names = [n for n, _ in m_new.named_parameters()]
    x, y = torch.from_numpy(d["xa"]).cuda(), torch.from_numpy(d["y"]).cuda()
    gen = torch.Generator(device="cuda").manual_seed(3)
    for step in range(6):
        xs = x if step == 0 else (x + 0.5 * torch.randn(x.shape, generator=gen, device="cuda")).clamp(-7.5, 7.5).mul(2).round().div(2)
        l_ref = train_step(m_ref, crit, opt, xs, y, T)
        l_new = tr.step(xs, y)
        # step 0 starts from identical weights; later steps see weights that differ in the last bits, and a QMS rounding
        # boundary may then fall differently for a message or two (the decoder is discontinuous there)
        tol = 5e-6 if step == 0 else 3e-4
        assert abs(float(l_ref) - float(l_new)) < tol * max(1.0, abs(float(l_ref))), (step, float(l_ref), float(l_new))
        for (n, a), (_, b) in zip(m_ref.named_parameters(), m_new.named_parameters()):
            assert float((a.detach() - b.detach()).abs().max()) < 1e-5, (step, n)
    assert tr.steps_done == 6 and tr.last_grad_norm > 0.0
    # the optimiser kernel writes the weights through raw pointers: the decode-only paths must see them (they read the
    # parameters live, there is no cached copy to go stale)
    with torch.no_grad():
        vn_w, cn_w, ucn_w, compute_ucn, ucn_mix = m_new.fold_weights(list(range(T)), x.device)
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    dec = {DecoderType.SP: 0, DecoderType.MS: 1, DecoderType.QMS: 2}[m_new.decoding_type]
    want = torch.ops.nldpc.boosted_forward(x, vn_w, cn_w, ucn_w, m_new.conn_mat.graph_id(x.device), T, dec, int(m_new.decoder_qms_qbit),
                                           -20.0, 20.0, bool(compute_ucn), bool(ucn_mix), None, None, None, 0, False, 2, 0, False)[0]
    assert torch.equal(m_new.decode_soft_last(x).view(torch.int32), want.view(torch.int32))
    # names / shapes / state_dict keys are untouched by the flat-vector aliasing
    assert [n for n, _ in m_new.named_parameters()] == names
    assert set(m_new.state_dict().keys()) == set(m_ref.state_dict().keys())
    if tag == "cn2vn3_qms":
        # first step from the fixture state reproduces the reference's loss (tools/gen_golden.py)
        d2, m2 = _fresh(tag)
        l0 = FusedTrainer(m2, crit, T, graph=graph).step(x, y)
        assert abs(float(l0) - float(d2["loss64"])) < 2e-6 * max(1.0, abs(float(d2["loss64"])))


def test_fused_trainer_neural_decoder():
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from neural_ldpc_decoder_torch_b200.training import FusedTrainer
    from test_neural_gpu import make_model
    d = load_golden("train_neural_wimax")
    T, B = d["w"].shape[0], d["xa"].shape[0]
    crit = LDPCDecoderLoss(LossType.BCE, etha=float(d["etha"]))
    m_ref = make_model(d["basegraph"], int(d["Z"]), T, B, d["w"], d["b"])
    m_new = make_model(d["basegraph"], int(d["Z"]), T, B, d["w"], d["b"])
    x = torch.from_numpy(d["xa"]).cuda()
    y = torch.zeros(B, m_ref.N * m_ref.Z, device="cuda")
    opt = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
    tr = FusedTrainer(m_new, crit, T, clamp=(-float("inf"), float("inf")))
    for _ in range(3):
        opt.zero_grad()
        loss = crit(m_ref(x), y, coeff_param=list(range(T)))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 1.0)
        opt.step()
        l_new = tr.step(x, y)
        assert abs(float(loss) - float(l_new)) < 5e-6 * max(1.0, abs(float(loss)))
    for a, b in zip(m_ref.parameters(), m_new.parameters()):
        assert float((a.detach() - b.detach()).abs().max()) < 1e-5


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_data_parallel_nccl(graph):
    """2 ranks over NCCL on shards of one batch follow a single process on the whole batch (tools/check_fused_trainer_ddp.py);
    graph=True: the whole step INCLUDING the NCCL all-reduce of the flat gradient replayed from a CUDA graph"""
    import json
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if graph else "29532", os.path.join(root, "tools", "check_fused_trainer_ddp.py")] + (["--graph"] if graph else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1500:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["ranks_identical"] and line["weights_moved_by"] > 1e-3


def test_fused_trainer_frozen_params_lr_schedule_and_state_dict():
    """(1) fixed_iterative_nodes_init_weight > 0: the frozen parameters' gradients enter the clipping norm like in
    clip_grad_norm_(model.parameters()) (train/...py:291) and no frozen weight moves; (2) set_lr reaches a CAPTURED step;
    (3) state_dict / load_state_dict resume the Adam state."""
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    from neural_ldpc_decoder_torch_b200.training import FusedTrainer, train_step
    d = load_golden("train_boosted_cn2vn3_qms")
    T, B = int(d["T"]), d["xa"].shape[0]
    x, y = torch.from_numpy(d["xa"]).cuda(), torch.from_numpy(d["y"]).cuda()
    crit = LDPCDecoderLoss(LossType.BCE, etha=1.0)

    def make():
        cm = ConnectingMatrixTorch(ConnectingMatrix(Z=int(d["Z"]), basegraph=d["basegraph"]), device=torch.device("cuda"))
        m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(2, 0, 3),
                                     fixed_iterative_nodes_init_weight=2).cuda()
        m.store_llr = "none"
        with torch.no_grad():
            for n, p in m.named_parameters():
                p.copy_(torch.from_numpy(d["param_" + n]))
        return m

    m_ref, m_new = make(), make()
    assert len(m_ref.get_trainable_parameters()) < len(list(m_ref.parameters()))
    opt = torch.optim.Adam(m_ref.get_trainable_parameters(), lr=1e-2)
    tr = FusedTrainer(m_new, crit, T, lr=1e-2, max_grad_norm=0.01)          # a norm bound that clips: the coefficient matters
    assert len(tr.frozen) == 4
    for step in range(3):
        train_step(m_ref, crit, opt, x, y, T, max_grad_norm=0.01)
        tr.step(x, y)
        # step 0 starts from identical weights: the clipped gradients (norm over ALL parameters) and the update must agree
        # closely; later steps see weights that differ in the last bits, which Adam's normalisation amplifies for components
        # whose gradient is nearly zero and which may move a QMS rounding boundary
        wtol, gtol = (2e-6, 3e-5) if step == 0 else (5e-4, None)
        for (n, a), (_, b) in zip(m_ref.named_parameters(), m_new.named_parameters()):
            assert float((a.detach() - b.detach()).abs().max()) < wtol, (step, n)
            if gtol is not None:
                assert float((a.grad - b.grad).abs().max()) <= gtol * max(float(a.grad.abs().max()), 1e-12), (step, n)   # clipped grads
        if step == 0:
            # the norm the kernel reports is the one clip_grad_norm_(model.parameters()) computes: frozen parameters included
            total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m_ref.parameters()))      # (after clipping: == max_norm)
            assert abs(float(total) - 0.01) < 1e-6 and tr.last_grad_norm > 0.01
    frozen0 = {n: torch.from_numpy(d["param_" + n]).cuda() for n, _ in m_new.named_parameters() if n.endswith(("_0", "_1"))}
    assert all(torch.equal(dict(m_new.named_parameters())[n].detach(), v) for n, v in frozen0.items())

    # (3) resume: a new trainer loaded with the state continues exactly like the old one
    sd = tr.state_dict()
    m_res = make()
    m_res.load_state_dict(m_new.state_dict())
    tr2 = FusedTrainer(m_res, crit, T, lr=123.0, max_grad_norm=0.01)
    tr2.load_state_dict(sd)
    tr.step(x, y)
    tr2.step(x, y)
    # (not bit-for-bit: the backward sweep's warps add their partial sums into the per-CTA totals with shared-memory atomics,
    # so the fp32 summation order of a gradient varies from run to run; Adam's normalisation amplifies last-bit differences
    # of near-zero components exactly as between the torch-op loop and the fused step above)
    for a, b in zip(m_new.parameters(), m_res.parameters()):
        assert float((a.detach() - b.detach()).abs().max()) < 5e-4
    assert tr2.steps_done == tr.steps_done == 4

    # (2) a captured step follows set_lr
    m_g = make()
    trg = FusedTrainer(m_g, crit, T, lr=1e-2, graph=True)
    trg.step(x, y)
    before = torch.cat([p.detach().reshape(-1).clone() for p in m_g.get_trainable_parameters()])
    trg.set_lr(0.0)
    trg.step(x, y)
    after = torch.cat([p.detach().reshape(-1) for p in m_g.get_trainable_parameters()])
    assert torch.equal(before, after)                                       # lr = 0: the replayed step moves nothing
    trg.set_lr(1e-2)
    trg.step(x, y)
    assert not torch.equal(before, torch.cat([p.detach().reshape(-1) for p in m_g.get_trainable_parameters()]))
