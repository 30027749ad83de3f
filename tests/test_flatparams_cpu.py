"""Host logic of the live-parameter decode path (_flatparams.py): the modules keep their parameters as slices of one flat
vector; the decode-only paths read [T, .] rows from it live (strided view / gather through a cached index map) and must
always equal what fold_weights / torch.stack produce — after in-place updates through .data, after .to(), after a broken
layout.  CPU tensors: this exercises the index logic only, no kernel is called."""
import copy

import pytest
import torch


def _boosted(graphs, sharing, T=6):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    bg, Z = graphs["wimax"]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cpu"))
    m = BoostedNeuralLDPCDecoder(T, 2, cm, node_weight_sharing_config=NodeWeightSharingConfig(*sharing))
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.rand_like(p))
    return m


def _same(a, b):
    for x, y in zip(a, b):
        if isinstance(y, torch.Tensor):
            assert torch.equal(x, y.detach())
        else:
            assert (x is None and y is None) or x == y, (x, y)


@pytest.mark.parametrize("sharing", [(3, 0, 3), (1, 0, 2), (2, 2, 3), (1, 1, 0), (3, 3, 3), (2, 1, 2), (0, 0, 2), (3, 0, 0), (0, 0, 0)])
def test_boosted_live_fold_equals_fold_weights(sharing, graphs):
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import _folded_live
    cpu, T = torch.device("cpu"), 6
    m = _boosted(graphs, sharing, T)
    ps = list(m.parameters())
    _same(_folded_live(m, T, cpu), m.fold_weights(list(range(T)), cpu))
    if ps:
        assert m.__dict__["_fold_index"][(cpu, T)][1] is not None, "flat layout not used"
        assert len({p.untyped_storage().data_ptr() for p in ps}) == 1          # one storage for every parameter
        ps[0].data.mul_(0.5)                                # no Tensor._version bump: must still be seen
        ps[-1].data.clamp_(0.2, 0.4)
        _same(_folded_live(m, T, cpu), m.fold_weights(list(range(T)), cpu))
        _same(_folded_live(m, 3, cpu), m.fold_weights(list(range(3)), cpu))     # fewer iterations: own index map
        m2 = copy.deepcopy(m).to(torch.float32).cpu()       # _apply re-establishes the layout
        assert len({p.untyped_storage().data_ptr() for p in m2.parameters()}) == 1
        _same(_folded_live(m2, T, cpu), m.fold_weights(list(range(T)), cpu))
        ps[0].data = torch.full_like(ps[0].data, 0.3)       # broken layout: plain fold_weights, still right
        _same(_folded_live(m, T, cpu), m.fold_weights(list(range(T)), cpu))
        assert m.__dict__["_fold_index"][(cpu, T)][1] is None
        names = [n for n, _ in m.named_parameters()]
        assert names == [n for n, _ in m2.named_parameters()]


def test_neural_rows_are_live_views(graphs):
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
    bg, Z = graphs["wimax"]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg), device=torch.device("cpu"))
    m = NeuralLDPCDecoder(5, 4, cm)
    cpu = torch.device("cpu")
    w, b = m._stacked_nograd(cpu)
    assert m.__dict__["_rows_view"][2] is not None and w.shape == (5, m.weights_var[0].numel())
    assert w.data_ptr() == m.weights_var[0].data_ptr() and b.data_ptr() == m.biases_var[0].data_ptr()      # views, not copies
    m.weights_var[3].data.fill_(0.25)
    with torch.no_grad():
        m.biases_var[1].add_(1.5)
    w2, b2 = m._stacked_nograd(cpu)
    ws, bs = m._stacked()
    assert torch.equal(w2, ws.detach()) and torch.equal(b2, bs.detach())
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    w3, _ = m._stacked_nograd(cpu)
    assert torch.equal(w3, m._stacked()[0].detach())
    m.weights_var[0].data = torch.zeros_like(m.weights_var[0].data)         # broken layout -> stacked copy
    w4, _ = m._stacked_nograd(cpu)
    assert m.__dict__["_rows_view"][2] is None and torch.equal(w4, m._stacked()[0].detach())
    sd = m.state_dict()
    assert "weights_var.0" in sd and "biases_var.4" in sd
