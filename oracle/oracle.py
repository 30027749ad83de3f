"""ctypes front-end of oracle/nldpc_oracle.c (see that file's header for the reference citations).

TEST INFRASTRUCTURE: the sparse CPU restatement of NeuralLDPCDecoder.forward
(/root/reference/src/neural_ldpc_decoder/NeuralLDPCDecoder.py:44-100) and of
BoostedNeuralLDPCDecoder.forward (…/BoostedNeuralLDPCDecoder.py:260-538).
Parity pinning: tests/golden/*.npz were produced by the live reference (tools/gen_golden.py)
and tests/test_oracle_golden.py checks this oracle against them bit-for-bit.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnldpc_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.nldpc_oracle_neural_forward.restype = ctypes.c_int
        L.nldpc_oracle_neural_forward.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  _f32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int, _f32p]
        L.nldpc_oracle_neural_forward_last.restype = ctypes.c_int
        L.nldpc_oracle_neural_forward_last.argtypes = L.nldpc_oracle_neural_forward.argtypes
        L.nldpc_oracle_boosted_step.restype = ctypes.c_int
        L.nldpc_oracle_boosted_step.argtypes = [_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_float, ctypes.c_float,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                _f32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_int]
        L.nldpc_oracle_pack_hard.restype = None
        L.nldpc_oracle_pack_hard.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _u8p]
        L.nldpc_oracle_quantize.restype = None
        L.nldpc_oracle_quantize.argtypes = [_f32p, _f32p, ctypes.c_long, ctypes.c_int]
        _lib = L
    return _lib


def _bg(basegraph):
    bg = np.ascontiguousarray(np.asarray(basegraph), dtype=np.int32)
    assert bg.ndim == 2
    return bg


def neural_forward(basegraph, Z, xa, w, b):
    """xa [B,N,Z] f32, w/b [T,E] f32 (row-major edge order) -> out [T,B,N*Z] f32."""
    bg = _bg(basegraph)
    M, N = bg.shape
    xa = np.ascontiguousarray(xa, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    B, T = xa.shape[0], w.shape[0]
    assert xa.shape == (B, N, Z) and w.shape == b.shape and w.shape[1] == int((bg != -1).sum())
    out = np.empty((T, B, N * Z), dtype=np.float32)
    rc = lib().nldpc_oracle_neural_forward(bg, M, N, Z, xa, w, b, B, T, out)
    if rc:
        raise RuntimeError(f"oracle neural_forward failed rc={rc}")
    return out


def neural_forward_last(basegraph, Z, xa, w, b):
    """neural_forward keeping only the last iteration: -> out [B,N*Z] f32 (full-size parity runs: T times less memory)."""
    bg = _bg(basegraph)
    M, N = bg.shape
    xa = np.ascontiguousarray(xa, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    B, T = xa.shape[0], w.shape[0]
    assert xa.shape == (B, N, Z) and w.shape == b.shape and w.shape[1] == int((bg != -1).sum())
    out = np.empty((B, N * Z), dtype=np.float32)
    rc = lib().nldpc_oracle_neural_forward_last(bg, M, N, Z, xa, w, b, B, T, out)
    if rc:
        raise RuntimeError(f"oracle neural_forward_last failed rc={rc}")
    return out


def _opt(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def boosted_step(basegraph, Z, decoder_type, qbit, llr_range, xin, xo, llr_in, vn_w=None, cn_w=None, ucn_w=None,
                 compute_ucn=False, ucn_mix=False, ucn_app=None):
    """One iteration of the Boosted loop body. xin/xo [B,N,Z] are updated IN PLACE (compounding VN weight,
    re-quantisation). llr_in [B,E,Z]. Returns (llr_out [B,E,Z], out [B,N*Z])."""
    bg = _bg(basegraph)
    M, N = bg.shape
    B = xin.shape[0]
    E = int((bg != -1).sum())
    assert xin.dtype == np.float32 and xin.flags.c_contiguous and xo.dtype == np.float32 and xo.flags.c_contiguous
    llr_in = np.ascontiguousarray(llr_in, dtype=np.float32)
    assert llr_in.shape == (B, E, Z)
    llr_out = np.empty((B, E, Z), dtype=np.float32)
    out = np.empty((B, N * Z), dtype=np.float32)
    k1, p1 = _opt(vn_w)
    k2, p2 = _opt(cn_w)
    k3, p3 = _opt(ucn_w)
    k4, p4 = _opt(ucn_app)
    rc = lib().nldpc_oracle_boosted_step(bg, M, N, Z, int(decoder_type), int(qbit), float(llr_range[0]), float(llr_range[1]),
                                         p1, p2, p3, int(bool(compute_ucn)), int(bool(ucn_mix)), p4,
                                         xin, xo, llr_in, llr_out, out, B)
    if rc:
        raise RuntimeError(f"oracle boosted_step failed rc={rc}")
    return llr_out, out


def boosted_forward(basegraph, Z, xa, T, decoder_type=2, qbit=5, llr_range=(-20.0, 20.0),
                    vn_w=None, cn_w=None, ucn_w=None, compute_ucn=False, ucn_mix=False, return_llr=False):
    """Full T-iteration Boosted forward from a zero state (target_iter=None semantics).
    vn_w [T,N] | None (only sharing types 2/3 multiply), cn_w/ucn_w [T,E] | None: sharing types folded by the caller.
    Returns out [T,B,N*Z] (and the c2v history [T+1,B,E,Z] if return_llr)."""
    bg = _bg(basegraph)
    xa = np.ascontiguousarray(xa, dtype=np.float32)
    B, N = xa.shape[0], xa.shape[1]
    E = int((bg != -1).sum())
    xin, xo = xa.copy(), xa.copy()
    llr = np.zeros((B, E, Z), dtype=np.float32)
    outs, llrs = [], [llr]
    for t in range(T):
        llr, out = boosted_step(bg, Z, decoder_type, qbit, llr_range, xin, xo, llr,
                                None if vn_w is None else vn_w[t], None if cn_w is None else cn_w[t],
                                None if ucn_w is None else ucn_w[t], compute_ucn, ucn_mix,
                                None if t == 0 else outs[-1])
        outs.append(out)
        llrs.append(llr)
    out = np.stack(outs)
    return (out, np.stack(llrs)) if return_llr else out


def pack_hard(out):
    """[B,NZ] f32 -> [B,ceil(NZ/8)] u8, bit = (out < 0), little-endian bit order (Functions.py:90)."""
    out = np.ascontiguousarray(out, dtype=np.float32)
    B, NZ = out.shape
    packed = np.empty((B, (NZ + 7) // 8), dtype=np.uint8)
    lib().nldpc_oracle_pack_hard(out, B, NZ, packed)
    return packed


def quantize(x, qbit):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().nldpc_oracle_quantize(x.reshape(-1), y.reshape(-1), x.size, int(qbit))
    return y


def count_errors(soft, y):
    """[T,B,NZ] f32 iteration outputs, [B,NZ] f32 labels -> int64 [2,T]: row 0 = positions where the decided bit
    ((out < 0) as 0.0/1.0, Functions.py:90) differs from the label (:93-94), row 1 = codewords with at least one (:98-99)."""
    soft = np.asarray(soft, dtype=np.float32)
    y = np.asarray(y, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        wrong = (soft < 0).astype(np.float32) != y[None]
    per = wrong.sum(axis=2, dtype=np.int64)
    return np.stack([per.sum(axis=1, dtype=np.int64), (per > 0).sum(axis=1, dtype=np.int64)]).astype(np.int64)


def count_errors_packed(hard, n_bits, y_packed=None):
    """the same on packed decisions [T,B,ceil(n_bits/8)] u8 (pack_hard layout); y_packed None = all-zero codeword"""
    hard = np.asarray(hard, dtype=np.uint8)
    bits = np.unpackbits(hard, axis=-1, bitorder="little")[..., :n_bits]
    if y_packed is not None:
        bits = bits ^ np.unpackbits(np.asarray(y_packed, dtype=np.uint8), axis=-1, bitorder="little")[None, ..., :n_bits]
    per = bits.sum(axis=2, dtype=np.int64)
    return np.stack([per.sum(axis=1, dtype=np.int64), (per > 0).sum(axis=1, dtype=np.int64)]).astype(np.int64)
