"""GPU: gradients of the CUDA backward kernels against the reference's autograd (golden fixtures from tools/gen_golden.py)
— Neural (weights + biases) and Boosted (train/train_BoostedNeuralLDPCDecoder.py:270-294 step: forward on all iterations,
multi-iteration BCE, backward).  fp32 tolerance: the kernel accumulates batch sums with atomics."""
import numpy as np
import pytest
import torch

from boosted_util import build_module
from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("code", ["bg2", "wimax"])
def test_neural_gradients_match_reference_autograd(code):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from test_neural_gpu import make_model
    d = load_golden(f"train_neural_{code}")
    T, B = d["w"].shape[0], d["xa"].shape[0]
    m = make_model(d["basegraph"], int(d["Z"]), T, B, d["w"], d["b"])
    outs = m(torch.from_numpy(d["xa"]).cuda())
    y = torch.zeros(B, outs[0].shape[1], device="cuda")
    loss = LDPCDecoderLoss(LossType.BCE, etha=float(d["etha"]))(outs, y, coeff_param=list(range(T)))
    assert abs(loss.item() - float(d["loss"])) < 1e-6 * max(1.0, abs(float(d["loss"])))
    loss.backward()
    gw = np.stack([p.grad.cpu().numpy() for p in m.weights_var])
    gb = np.stack([p.grad.cpu().numpy() for p in m.biases_var])
    assert rel_err(gw, d["grad_w"]) < 2e-5, rel_err(gw, d["grad_w"])
    assert rel_err(gb, d["grad_b"]) < 2e-5, rel_err(gb, d["grad_b"])
    # back-propagating from ONE list element (per-iteration optimisers, test_NeuralLDPCDecoder.py:104-109) also works
    for p in m.parameters():
        p.grad = None
    outs = m(torch.from_numpy(d["xa"]).cuda())
    torch.nn.functional.binary_cross_entropy_with_logits(outs[2], y).backward()
    assert all(float(m.weights_var[t].grad.abs().sum()) == 0.0 for t in range(3, T))
    assert float(m.weights_var[2].grad.abs().sum()) > 0.0


@pytest.mark.parametrize("tag", ["d4", "cn1vn2_ms", "cn2vn3_qms"])
def test_boosted_train_step_matches_reference(tag):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    d = load_golden(f"train_boosted_{tag}")
    T = int(d["T"])
    m = build_module(d, device="cuda")
    outs = m(torch.from_numpy(d["xa"]).cuda(), target_iter=list(range(T)))
    loss = LDPCDecoderLoss(LossType.BCE, etha=float(d["etha"]))(outs, torch.from_numpy(d["y"]).cuda(), coeff_param=list(range(T)))
    assert abs(loss.item() - float(d["loss64"])) < 2e-6 * max(1.0, abs(float(d["loss64"])))
    loss.backward()
    for n, p in m.named_parameters():
        ref = d["grad_" + n]
        got = p.grad.cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(got - ref).max() < 3e-5 * scale + 1e-9, (n, got.reshape(-1)[:3], ref.reshape(-1)[:3])
    if tag == "d4":   # SURVEY.md Appendix D4 known answers
        assert abs(float(m.weight_CN_0.grad) - 0.02436903864145279) < 1e-6
        assert abs(float(m.weight_CN_1.grad) - 0.023220574483275414) < 1e-6
        gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())))
        assert abs(gn - 0.046869996935129166) < 1e-6


def test_fused_multi_iter_bce_matches_torch():
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    g = torch.Generator(device="cuda").manual_seed(1)
    T, B, NZ = 7, 33, 832
    base = (6.0 * torch.randn((T, B, NZ), generator=g, device="cuda")).requires_grad_(True)
    y = (torch.rand((B, NZ), generator=g, device="cuda") > 0.5).float()
    outs = list(base.unbind(0))
    crit = LDPCDecoderLoss(LossType.BCE, etha=1.3)
    fused = crit(outs, y, coeff_param=list(range(T)))
    fused.backward()
    gf = base.grad.clone()
    base.grad = None
    crit.fused = False
    plain = crit(list(base.unbind(0)), y, coeff_param=list(range(T)))
    plain.backward()
    assert abs(fused.item() - plain.item()) < 2e-6 * abs(plain.item())
    assert float((gf - base.grad).abs().max()) < 1e-6 * float(base.grad.abs().max()) + 1e-12
    # a list that is NOT a set of views of one tensor silently takes the per-iteration path
    crit.fused = True
    loose = [base[t].clone() for t in range(T)]
    assert abs(crit(loose, y, coeff_param=list(range(T))).item() - plain.item()) < 2e-6 * abs(plain.item())
