"""GPU: the drop-in module replays the recorded forward() calls of the live reference (tools/gen_golden_staging.py) and must
return the same tensors and leave the same public state (self.outputs, self.llr) after every call, bit for bit."""
import numpy as np
import pytest
import torch

from conftest import golden_json, load_golden
from staging_util import build_staging_module, call_args

pytestmark = pytest.mark.gpu
CASES = golden_json("staging_index.json")


def _run_calls(d, m, check_state=True):
    for k in range(int(d["n_calls"])):
        xa, target, fixed, fw = call_args(d, k)
        xin = [torch.from_numpy(a).cuda() for a in xa] if isinstance(xa, list) else torch.from_numpy(xa).cuda()
        fwt = None if fw is None else [torch.from_numpy(w).cuda() for w in fw]
        with torch.no_grad():
            ret = m(xin, target_iter=target, fixed_iter=fixed, fixed_iter_weight=fwt)
        got = ret.cpu().numpy()[None] if isinstance(ret, torch.Tensor) else np.stack([o.cpu().numpy() for o in ret])
        assert np.array_equal(got.view(np.uint32), d[f"c{k}_ret"].view(np.uint32)), (k, np.abs(got - d[f"c{k}_ret"]).max())
        if check_state:
            outs = np.stack([o.cpu().numpy() for o in m.outputs])
            llr = np.stack([l.cpu().numpy() for l in m.llr])
            assert np.array_equal(outs.view(np.uint32), d[f"c{k}_outputs"].view(np.uint32)), k
            assert np.array_equal(llr.view(np.uint32), d[f"c{k}_llr"].view(np.uint32)), k


@pytest.mark.parametrize("name", CASES)
def test_module_replays_reference_call_sequence(name):
    d = load_golden(name)
    m = build_staging_module(d, device="cuda")
    assert m.store_llr == "all"            # the default reproduces the reference's state
    _run_calls(d, m)


@pytest.mark.parametrize("name", [CASES[6]])
def test_store_last_keeps_staged_sequences_and_guards_gaps(name):
    """store_llr='last' (throughput setting): consecutive stages still chain; a stage whose entry state was skipped raises."""
    d = load_golden(name)                  # staged [0,1,2] -> [3..6] -> 7
    m = build_staging_module(d, device="cuda")
    m.store_llr = "last"
    _run_calls(d, m, check_state=False)
    xa, _, _, _ = call_args(d, 0)
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(xa).cuda(), target_iter=[2])       # llr[2] of the last run over iteration 1 was not stored
    m.store_llr = "none"
    with torch.no_grad():
        m(torch.from_numpy(xa).cuda())
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(xa).cuda(), target_iter=[5])


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("tag", ["bg2_qms303", "wimax_ms112", "bg2_qms100"])
def test_staged_training_backward_from_stored_state(tag, generic, monkeypatch):
    """train/train_BoostedNeuralLDPCDecoder.py:139-181, 270-294 with fixed_iter > 0: iterations [0, t0) under no_grad, then
    forward(target_iter=range(t0, T)) -> LDPCDecoderLoss -> backward.  The stored state (self.llr[t0], self.outputs[t0-1]) is
    a constant of the call; loss and parameter gradients must equal the reference's autograd (fixture: live reference)."""
    from boosted_util import build_module
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.LDPCDecoderLoss import LDPCDecoderLoss
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.LossType import LossType
    if generic:
        monkeypatch.setenv("NLDPC_FORCE_GENERIC", "1")
    d = load_golden("train_staged_" + tag)
    T, t0 = int(d["T"]), int(d["t0"])
    m = build_module(d, device="cuda")
    xa, y = torch.from_numpy(d["xa"]).cuda(), torch.from_numpy(d["y"]).cuda()
    with torch.no_grad():
        m(xa, target_iter=list(range(t0)))
    outs = m(xa, target_iter=list(range(t0, T)))
    got = np.stack([o.detach().cpu().numpy() for o in outs])
    assert np.array_equal(got.view(np.uint32), d["out"].view(np.uint32))
    loss = LDPCDecoderLoss(LossType.BCE, etha=float(d["etha"]))(outs, y, coeff_param=list(range(len(outs))))
    loss.backward()
    assert abs(float(loss) - float(d["loss"])) <= 2e-6 * max(1.0, abs(float(d["loss"])))
    want = {k[len("grad_"):]: d[k] for k in d.files if k.startswith("grad_")}
    have = {n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters() if p.grad is not None and float(p.grad.abs().sum()) > 0}
    assert set(want) >= set(have)                  # no gradient reaches the iterations that ran under no_grad
    for n, gref in want.items():
        g = dict(m.named_parameters())[n].grad
        assert g is not None, n
        scale = max(float(np.abs(gref).max()), 1e-12)
        assert float(np.abs(g.detach().cpu().numpy() - gref).max()) <= 3e-5 * scale, (n, g.reshape(-1)[:4], gref.reshape(-1)[:4])
