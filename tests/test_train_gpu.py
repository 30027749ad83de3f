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
    # the upstream gradient (a device scalar) is folded into the gradient kernel: scaled loss
    base.grad = None
    crit.fused = True
    (0.37 * crit(list(base.unbind(0)), y, coeff_param=list(range(T)))).backward()
    assert float((base.grad - 0.37 * gf).abs().max()) < 1e-6 * float(gf.abs().max()) + 1e-12
    base.grad = None
    # a list that is NOT a set of views of one tensor silently takes the per-iteration path
    crit.fused = True
    loose = [base[t].clone() for t in range(T)]
    assert abs(crit(loose, y, coeff_param=list(range(T))).item() - plain.item()) < 2e-6 * abs(plain.item())


def _grads(model_fn, x, y, T, generic):
    """gradients of one multi-iteration BCE step; `generic` forces the table-driven kernels (NLDPC_FORCE_GENERIC)"""
    import os
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    os.environ["NLDPC_FORCE_GENERIC"] = "1" if generic else "0"
    try:
        m = model_fn()
        outs = m(x) if not hasattr(m, "fold_weights") else m(x, target_iter=list(range(T)))
        loss = LDPCDecoderLoss(LossType.BCE, etha=1.0)(outs, y, coeff_param=list(range(T)))
        loss.backward()
        torch.cuda.synchronize()
        return {n: p.grad.detach().cpu().numpy().copy() for n, p in m.named_parameters() if p.grad is not None}
    finally:
        os.environ["NLDPC_FORCE_GENERIC"] = "0"


@pytest.mark.parametrize("code,B", [("nr_bg2_set0", 148 * 16 * 2 + 5), ("wimax_n576_r34", 37), ("nr_bg2_set0", 1)])
@pytest.mark.parametrize("kind", ["neural", "ms_cn", "qms_cn_vn", "qms_cn2_vn2"])
def test_specialised_backward_equals_table_driven(code, B, kind):
    """the graph-as-immediates backward sweep (nldpc_spec_backward.cuh) against the table-driven kernel on ragged batches:
    last tile partly empty, several tiles per CTA, a single codeword"""
    from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
    from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator
    bg, Z = load_basegraph(code)
    graph = TannerGraph(bg, Z)
    dev = torch.device("cuda")
    T = 4
    qbit = 5 if kind.startswith("qms") else None
    x, y = DeviceBatchGenerator(graph, [1.0, 2.0, 3.0], dev, seed=7, qms_qbit=qbit)(B)
    gen = torch.Generator().manual_seed(3)

    if kind == "neural":
        from test_neural_gpu import make_model
        w = (0.3 + torch.rand((T, graph.E), generator=gen)).numpy()
        b = (0.2 * torch.randn((T, graph.E), generator=gen)).numpy()
        model_fn = lambda: make_model(bg, Z, T, B, w, b)   # noqa: E731
    else:
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
        sharing = {"ms_cn": (1, 0, 0), "qms_cn_vn": (3, 0, 3), "qms_cn2_vn2": (2, 0, 2)}[kind]
        dec = DecoderType.MS if kind == "ms_cn" else DecoderType.QMS
        vals = {}

        def model_fn():
            cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
            m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=dec).to(dev)
            with torch.no_grad():
                for n, p in m.named_parameters():
                    if n not in vals:
                        vals[n] = 0.6 + 0.6 * torch.rand(p.shape, generator=gen)
                    p.copy_(vals[n].to(dev))
            return m

    spec = _grads(model_fn, x, y, T, generic=False)
    tabl = _grads(model_fn, x, y, T, generic=True)
    assert spec.keys() == tabl.keys() and len(spec) > 0
    for n in spec:
        scale = max(float(np.abs(tabl[n]).max()), 1e-12)
        assert float(np.abs(spec[n] - tabl[n]).max()) <= 3e-5 * scale, (n, spec[n].reshape(-1)[:4], tabl[n].reshape(-1)[:4])


def test_two_streams_do_not_share_constant_arena_ranges():
    """decode + gradient launches issued back to back on two CUDA streams with DIFFERENT weights: each launch packs its weights
    into its own range of the constant arena (ring allocator with per-range events), so results must equal the serial runs"""
    from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
    from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator
    from test_neural_gpu import make_model
    bg, Z = load_basegraph("nr_bg2_set0")
    graph = TannerGraph(bg, Z)
    dev = torch.device("cuda")
    T, B = 6, 3000
    x, y = DeviceBatchGenerator(graph, [1.5, 3.0], dev, seed=11)(B)
    gen = torch.Generator().manual_seed(5)
    ws = [((0.3 + torch.rand((T, graph.E), generator=gen)).numpy(), (0.2 * torch.randn((T, graph.E), generator=gen)).numpy()) for _ in range(2)]
    models = [make_model(bg, Z, T, B, w, b) for w, b in ws]

    def run(m):
        for p in m.parameters():
            p.grad = None
        outs = m(x)
        torch.nn.functional.binary_cross_entropy_with_logits(-outs[-1], y).backward()
        return outs[-1].detach().clone(), torch.stack([p.grad for p in m.weights_var]).clone()

    serial = [run(m) for m in models]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    conc = [None, None]
    for rep in range(5):
        for i in (0, 1):
            streams[i].wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(streams[i]):
                conc[i] = run(models[i])
        torch.cuda.synchronize()
        for i in (0, 1):
            assert torch.equal(conc[i][0], serial[i][0])
            scale = float(serial[i][1].abs().max())
            assert float((conc[i][1] - serial[i][1]).abs().max()) <= 3e-5 * scale


# ---- fused training forward (forward + multi-iteration BCE + dL/dout in one launch) ---------------------------------------
@pytest.mark.parametrize("tag", ["d4", "cn1vn2_ms", "cn2vn3_qms"])
def test_fused_bce_loss_matches_reference_autograd(tag):
    """model.fused_bce_loss(x, y) == LDPCDecoderLoss(BCE)(model(x), y): loss and gradients against the REFERENCE's autograd
    goldens (same fixtures and tolerances as the two-call form above; fp32, tolerance 3e-5 of the largest gradient)"""
    d = load_golden(f"train_boosted_{tag}")
    T = int(d["T"])
    m = build_module(d, device="cuda")
    x, y = torch.from_numpy(d["xa"]).cuda(), torch.from_numpy(d["y"]).cuda()
    loss = m.fused_bce_loss(x, y, etha=float(d["etha"]), coeff_param=list(range(T)))
    assert loss is not None, "the built-in codes with CN weights are covered by the fused path"
    assert abs(loss.item() - float(d["loss64"])) < 5e-6 * max(1.0, abs(float(d["loss64"])))
    loss.backward()
    for n, p in m.named_parameters():
        ref = d["grad_" + n]
        got = p.grad.cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-6)
        assert np.abs(got - ref).max() < 3e-5 * scale + 1e-9, (n, got.reshape(-1)[:3], ref.reshape(-1)[:3])
    # an upstream factor reaches the weight gradients
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    for p in m.parameters():
        p.grad = None
    (0.25 * m.fused_bce_loss(x, y, etha=float(d["etha"]), coeff_param=list(range(T)))).backward()
    for n, p in m.named_parameters():
        assert float((p.grad - 0.25 * g1[n]).abs().max()) <= 2e-6 * float(g1[n].abs().max()) + 1e-12


@pytest.mark.parametrize("code,B", [("nr_bg2_set0", 148 * 16 + 3), ("wimax_n576_r34", 37), ("nr_bg2_set0", 1)])
@pytest.mark.parametrize("kind", ["ms_cn", "qms_cn_vn", "qms_cn2_vn2"])
def test_fused_bce_loss_equals_two_call_form(code, B, kind):
    """ragged batches, random (non-zero) codeword labels, etha != 1: the fused launch against forward() + LDPCDecoderLoss +
    backward() of the same module (which the tests above pin to the reference)"""
    from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    from neural_ldpc_decoder_torch_b200.training import DeviceBatchGenerator
    bg, Z = load_basegraph(code)
    graph = TannerGraph(bg, Z)
    dev = torch.device("cuda")
    T = 5
    qbit = 5 if kind.startswith("qms") else None
    x, y = DeviceBatchGenerator(graph, [1.0, 2.0, 3.0], dev, seed=11, qms_qbit=qbit)(B)
    sharing = {"ms_cn": (1, 0, 0), "qms_cn_vn": (3, 0, 3), "qms_cn2_vn2": (2, 0, 2)}[kind]
    dec = DecoderType.MS if kind == "ms_cn" else DecoderType.QMS
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    m = BoostedNeuralLDPCDecoder(T, B, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=dec).to(dev)
    m.store_llr = "none"
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_((0.6 + 0.8 * torch.rand(p.shape, generator=gen)).to(dev))
    crit = LDPCDecoderLoss(LossType.BCE, etha=1.2)
    ref_loss = crit(m(x, target_iter=list(range(T))), y, coeff_param=list(range(T)))
    ref_loss.backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters()}
    for p in m.parameters():
        p.grad = None
    loss = m.fused_bce_loss(x, y, etha=1.2, coeff_param=list(range(T)))
    assert loss is not None
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 5e-6 * max(1.0, abs(ref_loss.item()))
    for n, p in m.named_parameters():
        scale = max(float(ref[n].abs().max()), 1e-9)
        assert float((p.grad - ref[n]).abs().max()) < 3e-5 * scale, (n, float((p.grad - ref[n]).abs().max()), scale)


def test_fused_bce_loss_declines_what_it_does_not_cover():
    from neural_ldpc_decoder_torch_b200 import load_basegraph
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = load_basegraph("wimax_n576_r34")
    dev = torch.device("cuda")
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=dev)
    x = torch.randn(4, bg.shape[1], Z, device=dev)
    y = torch.zeros(4, bg.shape[1] * Z, device=dev)
    for sharing, dec in (((3, 3, 0), DecoderType.QMS), ((0, 0, 3), DecoderType.QMS), ((3, 0, 0), DecoderType.SP)):
        m = BoostedNeuralLDPCDecoder(3, 4, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing), decoding_type=dec).to(dev)
        assert m.fused_bce_loss(x, y) is None


def test_pack_labels_matches_numpy_packbits():
    from neural_ldpc_decoder_torch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    for B, NZ in ((5, 832), (3, 576), (2, 13)):
        y = (torch.rand((B, NZ), generator=g, device="cuda") > 0.5).float()
        got = ops.pack_labels(y).cpu().numpy()
        ref = np.packbits(y.cpu().numpy().astype(np.uint8), axis=1, bitorder="little")
        assert np.array_equal(got, ref)
