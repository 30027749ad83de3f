"""CPU: host-side logic — graph tables, drop-in mirror classes, state_dict layout, data generator, C-ABI symbols."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def test_builtin_graph_constants(graphs):
    from neural_ldpc_decoder_torch_b200 import TannerGraph
    bg, Z = graphs["bg2"]
    g = TannerGraph(bg, Z)
    assert (g.M, g.N, g.Z, g.E) == (42, 52, 16, 197)
    assert list(g.col_deg[:14]) == [22, 23, 10, 5, 5, 14, 7, 13, 6, 8, 9, 16, 9, 12] and (g.col_deg[14:] == 1).all()
    assert list(g.row_deg[:6]) == [8, 10, 8, 10, 4, 6]
    bg, Z = graphs["wimax"]
    g = TannerGraph(bg, Z)
    assert (g.M, g.N, g.Z, g.E) == (6, 24, 24, 88)
    assert list(g.row_deg) == [14, 15, 15, 15, 14, 15]


def test_lifted_H_and_systematic_generator(graphs):
    from neural_ldpc_decoder_torch_b200 import TannerGraph
    for code in ("bg2", "wimax"):
        bg, Z = graphs[code]
        g = TannerGraph(bg, Z)
        H, G = g.lifted_H(), g.systematic_generator()
        assert G.shape == ((g.N - g.M) * Z, g.N * Z)
        assert (H.astype(np.int64) @ G.T.astype(np.int64) % 2).sum() == 0
        assert np.array_equal(G[:, :G.shape[0]], np.eye(G.shape[0], dtype=np.uint8))
    # SHA of the BG2 z16 generator == the reference's resources/gen_matrix_bg2_z16.txt (checked in the build container)
    bg, Z = graphs["bg2"]
    G = TannerGraph(bg, Z).systematic_generator()
    assert hashlib.sha256(G.astype(np.uint8).tobytes()).hexdigest()[:16] == GEN_BG2_SHA


GEN_BG2_SHA = "64e8cff5430b2c4a"


def test_datagen_reproduces_reference_stream(graphs):
    """the golden xa was produced by the reference's AWGNPassedDatagen (seeds 2042/1074, 2 dB)"""
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import AWGNPassedDatagen
    for code, name in (("bg2", "neural_bg2_init"), ("wimax", "neural_wimax_init")):
        bg, Z = graphs[code]
        M, N = bg.shape
        d = load_golden(name)
        gen = np.zeros(((N - M) * Z, N * Z), dtype=np.int64)
        dg = AWGNPassedDatagen(N=N, M=M, snr_db=np.array([2.0]), awgn_noise_seed=2042, wordgen_random_seed=1074, gen_matrix=gen)
        x, y = dg(word_length=8, Z=Z, is_y_all_zero=True)
        assert x[0].dtype == np.float32 and np.array_equal(np.reshape(x[0], [8, N, Z]), d["xa"])
        assert (y[0] == 0).all()


def test_neural_module_api_and_state_dict(graphs):
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
    bg, Z = graphs["wimax"]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg))
    m = NeuralLDPCDecoder(3, 4, cm)
    E = 88
    assert (m.N, m.M, m.Z, int(m.sum_edge)) == (24, 6, 24, E)
    assert len(m.weights_var) == 3 and all(torch.equal(p, torch.full((E,), 0.5)) for p in m.weights_var)
    assert all(torch.equal(p, torch.zeros(E)) for p in m.biases_var)
    sd = m.state_dict()
    assert list(sd.keys()) == ["W_odd2even", "W_skipconn2even", "W_even2odd", "W_output", "Lift_Matrix1", "Lift_Matrix2",
                               "weights_var.0", "weights_var.1", "weights_var.2", "biases_var.0", "biases_var.1", "biases_var.2"]
    assert tuple(sd["W_odd2even"].shape) == (E, E) and tuple(sd["W_skipconn2even"].shape) == (24, E)
    assert tuple(sd["Lift_Matrix1"].shape) == (E * Z, E * Z) and tuple(sd["W_output"].shape) == (E, 24)
    assert float(sd["Lift_Matrix1"].sum()) == E * Z and float(sd["W_output"].sum()) == E
    m2 = NeuralLDPCDecoder(3, 4, cm)
    with torch.no_grad():
        m.weights_var[1].fill_(0.25)
    m2.load_state_dict(m.state_dict())                      # strict, accepts (and drops) the dense buffers
    assert torch.equal(m2.weights_var[1], torch.full((E,), 0.25))
    # buffers are readable attributes like in the reference
    assert tuple(m.W_even2odd.shape) == (E, E)


def test_dropin_import_names():
    import neural_ldpc_decoder_torch_b200 as nl
    nl.install_dropin()
    import neural_ldpc_decoder as N
    import neural_ldpc_decoder.NeuralLDPCDecoder as cls    # the reference's tests import the class this way
    assert isinstance(cls, type) and cls is N.NeuralLDPCDecoder
    for name in ("AWGNPassedDatagen", "ConnectingMatrix", "ConnectingMatrixTorch", "NeuralLDPCDecoder"):
        assert hasattr(N, name)


def test_c_abi_exports_every_declared_symbol():
    """the shared library loads without a GPU and exports every function include/nldpc.h declares"""
    so = os.path.join(ROOT, "neural_ldpc_decoder_torch_b200", "libnldpc_b200.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "nldpc.h")).read()
    names = sorted(set(re.findall(r"\b(nldpc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nldpc.h but not exported"
    assert lib.nldpc_abi_version() == 4


def test_product_package_never_imports_oracle():
    """the product may mention the oracle in comments, but must not import / include / dlopen it"""
    pkg = os.path.join(ROOT, "neural_ldpc_decoder_torch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                for line in open(path):
                    code = line.split("#")[0]
                    assert not re.search(r"\b(import|from)\s+oracle\b", code), (f, line)
                    assert "libnldpc_oracle" not in code, (f, line)
            elif f.endswith((".cu", ".cuh", ".h")):
                for line in open(path):
                    assert not re.search(r"#\s*include.*oracle", line), (f, line)
                    assert "dlopen" not in line, (f, line)


def test_checkpoint_utils_roundtrip_and_reference_layout(tmp_path, graphs):
    from neural_ldpc_decoder_torch_b200.checkpoint_utils import CheckPointUtil, MetricsLogger
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
    bg, Z = graphs["wimax"]
    cm = ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg))
    m = NeuralLDPCDecoder(2, 4, cm)
    with torch.no_grad():
        m.weights_var[1].fill_(0.7)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    ck = CheckPointUtil(checkpoint_dir=str(tmp_path))
    path = ck.save("c.pth", m, optimizer=opt, epoch=3, metrics={"loss": 0.5, "ber_last_iter": 1e-3}, config={"T": 2})
    raw = torch.load(path)
    assert set(raw.keys()) == {"model_state_dict", "optimizer_state_dict", "epoch", "loss", "ber_last_iter", "config"}
    assert "Lift_Matrix1" in raw["model_state_dict"] and "weights_var.1" in raw["model_state_dict"]     # reference key set
    m2 = NeuralLDPCDecoder(2, 4, cm)
    got = ck.load("c.pth", m2, optimizer=torch.optim.Adam(m2.parameters()))
    assert got["epoch"] == 3 and torch.equal(m2.weights_var[1], m.weights_var[1])
    wpath = ck.save_weights("w_epoch_1", m, as_txt=True)
    assert wpath.endswith("w_epoch_1.pth") and os.path.exists(wpath)
    txt = os.path.join(str(tmp_path), "w_epoch_1_weights_txt")
    assert os.path.exists(os.path.join(txt, "weights_var_1.txt")) and not os.path.exists(os.path.join(txt, "Lift_Matrix1.txt"))
    assert np.allclose(np.loadtxt(os.path.join(txt, "weights_var_1.txt")), 0.7)
    log = MetricsLogger(log_dir=str(tmp_path))
    log.log(0, {"loss": 0.25, "ber_last_iter": 1.5e-3}, "c.pth", config={"T": 2})
    log.log(1, {"loss": 0.2, "ber_last_iter": 1.0e-3}, "c2.pth")
    lines = open(log.log_file).read().splitlines()
    assert lines[0].startswith("# Training started") and lines[2].startswith("# Columns: Epoch, Timestamp, loss, ber_last_iter")
    assert lines[-1].endswith("0.200000, 1.000000e-03, c2.pth")
    assert log.is_best(1e-3) and not log.is_best(2e-3) and log.is_best(5e-4)


def _sha_t(t):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32)).tobytes()).hexdigest()


@pytest.mark.parametrize("code", ["bg2", "wimax"])
@pytest.mark.parametrize("family", ["neural", "boosted"])
def test_dense_structure_matrices_equal_the_reference(code, family, graphs):
    """every dense matrix synthesised for state_dict() is byte-identical to the one the reference builds
    (neural ConnectingMatrix.py:68-140, boosted :82-163); SHA-256 fixtures from tools/gen_golden_checkpoint.py"""
    from conftest import golden_json
    want = golden_json("dense_matrix_sha.json")[f"{code}_{family}"]
    bg, Z = graphs[code]
    if family == "neural":
        from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
        m = NeuralLDPCDecoder(1, 2, ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg)))
    else:
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch
        from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
        m = BoostedNeuralLDPCDecoder(1, 2, ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg)))
    sd = m.state_dict()
    dense = [k for k in sd if k.startswith(("W_", "Lift_"))]
    assert dense == list(want.keys())                                   # same keys, same order
    for k in dense:
        assert list(sd[k].shape) == want[k]["shape"] and sd[k].dtype == torch.float32, k
        assert _sha_t(sd[k]) == want[k]["sha256"], k
        assert _sha_t(getattr(m, k)) == want[k]["sha256"], k           # ... and as the attribute the reference exposes


def test_reference_written_checkpoints_load_strictly(tmp_path):
    """checkpoints written by the REFERENCE's CheckPointUtil.save (fixtures: tools/gen_golden_checkpoint.py, toy code M=3 N=6
    Z=4) load through the mirror's strict load (CheckPointUtil.py:125-159), and the mirror's own state_dict() reproduces the
    reference's file key for key, value for value"""
    import shutil
    from conftest import GOLDEN
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder import ConnectingMatrix as BCM, ConnectingMatrixTorch as BCMT
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.BoostedNeuralLDPCDecoder import BoostedNeuralLDPCDecoder
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.DecoderType import DecoderType
    from neural_ldpc_decoder_torch_b200.boosted_neural_ldpc_decoder.struct.NodeWeightSharingConfig import NodeWeightSharingConfig
    from neural_ldpc_decoder_torch_b200.checkpoint_utils import CheckPointUtil
    from neural_ldpc_decoder_torch_b200.neural_ldpc_decoder import ConnectingMatrix, ConnectingMatrixTorch, NeuralLDPCDecoder
    toy = np.load(os.path.join(GOLDEN, "ref_ckpt_toy_graph.npz"))
    bg, Z = toy["basegraph"], int(toy["Z"])
    for f in ("ref_ckpt_neural_toy.pth", "ref_ckpt_boosted_toy.pth"):
        shutil.copy(os.path.join(GOLDEN, f), tmp_path / f)
    ck = CheckPointUtil(checkpoint_dir=str(tmp_path))

    m = NeuralLDPCDecoder(3, 2, ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg)))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    got = ck.load("ref_ckpt_neural_toy.pth", m, optimizer=opt)
    assert got["epoch"] == 7 and got["loss"] == 0.125 and got["config"] == {"T": 3, "code": "toy"}
    ref_sd = got["model_state_dict"]
    own_sd = m.state_dict()
    assert list(own_sd.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert own_sd[k].dtype == ref_sd[k].dtype and torch.equal(own_sd[k], ref_sd[k]), k
    assert len(opt.state_dict()["state"]) == 6 and opt.state_dict()["state"][0]["step"] == 1      # Adam moments came along

    mb = BoostedNeuralLDPCDecoder(4, 2, BCMT(BCM(Z=Z, basegraph=bg)), node_weight_sharing_config=NodeWeightSharingConfig(2, 2, 3),
                                  decoding_type=DecoderType.QMS)
    gotb = ck.load("ref_ckpt_boosted_toy.pth", mb, optimizer=torch.optim.Adam(mb.get_trainable_parameters()))
    ref_sd, own_sd = gotb["model_state_dict"], mb.state_dict()
    assert list(own_sd.keys()) == list(ref_sd.keys())
    for k in ref_sd:
        assert torch.equal(own_sd[k], ref_sd[k]), k
    assert gotb["epoch"] == 2

    # a checkpoint written for another graph must not load (the reference fails with a size mismatch)
    bg2 = bg.copy()
    bg2[0, 0] = 2                                                                                   # same shape, different shift
    other = NeuralLDPCDecoder(3, 2, ConnectingMatrixTorch(ConnectingMatrix(Z=Z, basegraph=bg2)))
    with pytest.raises(RuntimeError):
        ck.load("ref_ckpt_neural_toy.pth", other)
    wider = NeuralLDPCDecoder(3, 2, ConnectingMatrixTorch(ConnectingMatrix(Z=Z + 1, basegraph=bg)))
    with pytest.raises(RuntimeError):
        ck.load("ref_ckpt_neural_toy.pth", wider)


@pytest.mark.parametrize("code,header", [("nr_bg2_set0", "nldpc_graph_bg2z16.cuh"), ("wimax_n576_r34", "nldpc_graph_wimaxz24.cuh")])
def test_generated_weight_pair_layout(code, header):
    """csrc/generated/*.cuh: wb_pair_off() (the host table of the Neural {w, b} pair layout, tools/gen_kernels.py) is a
    permutation of an iteration's 2*E floats that stays inside each check's own 2*D floats, pairs are (w_a, w_b)(b_a, b_b) in
    two consecutive 8-byte entries, singles are (w, b) in one — what csrc/nldpc_spec_kernel.cuh WbPlan derives at compile time."""
    import os
    import re

    from neural_ldpc_decoder_torch_b200 import TannerGraph, load_basegraph
    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "..", "neural_ldpc_decoder_torch_b200", "csrc", "generated", header)).read()
    m = re.search(r"wb_pair_off\(\) \{\s*static const int t\[2 \* E\] = \{([^}]*)\}", src)
    assert m, "generated header has no wb_pair_off table (run tools/gen_kernels.py)"
    tab = [int(v) for v in m.group(1).split(",")]
    bg, Z = load_basegraph(code)
    g = TannerGraph(bg, Z)
    E = g.E
    assert len(tab) == 2 * E and sorted(tab) == list(range(2 * E))
    woff, boff = tab[:E], tab[E:]
    col_deg = np.bincount(np.asarray(g.ecol), minlength=g.N)
    for i in range(g.M):
        es = list(range(int(g.row_ptr[i]), int(g.row_ptr[i + 1])))
        inside = range(2 * es[0], 2 * (es[-1] + 1))
        assert all(woff[e] in inside and boff[e] in inside for e in es)
        stored = [e for e in es if col_deg[int(g.ecol[e])] >= 2]
        for a, b in zip(stored[0::2], stored[1::2]):
            assert woff[a] % 4 in (0, 2) and woff[a] % 2 == 0 and woff[b] == woff[a] + 1
            assert boff[a] == woff[a] + 2 and boff[b] == woff[a] + 3
        singles = ([stored[-1]] if len(stored) % 2 else []) + [e for e in es if col_deg[int(g.ecol[e])] < 2]
        for e in singles:
            assert woff[e] % 2 == 0 and boff[e] == woff[e] + 1
