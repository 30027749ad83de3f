"""torch.library custom ops over the C ABI (include/nldpc.h).  PyTorch is plumbing only: it owns the
device memory and the current stream; the arithmetic is in libnldpc_b200.so (hand-written sm_100a CUDA).

  nldpc::neural_forward(xa, w, b, graph_id) -> out [T, B, N*Z]       NeuralLDPCDecoder.py:54-98
  nldpc::neural_hard(xa, w, b, graph_id, all_iters) -> packed uint8  + Functions.py:90 predicate
  nldpc::neural_backward(xa, w, b, gout, graph_id) -> (gw, gb)       autograd of the above
"""
import ctypes
from typing import Optional

import torch

from . import _lib

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream(t):
    return _vp(torch.cuda.current_stream(t.device).cuda_stream)


def _aligned16(t):
    """contiguous, 16-byte aligned storage: the kernels stage xa with bulk TMA copies (cp.async.bulk needs 16-byte aligned
    sources); a contiguous view at an odd storage offset (flat[1:1+n].view(B, N, Z)) is copied once instead of faulting"""
    t = t.contiguous()
    return t.clone() if t.data_ptr() % 16 else t


def _check_cuda_f32(name, t):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise _lib.NldpcError(f"{name} is on {t.device}: the B200 decode path has no CPU fallback — move it to a CUDA device")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")


def _prep(xa, w, b, graph_id):
    g = _lib.graph_by_id(graph_id)
    for n, t in (("xa", xa), ("w", w), ("b", b)):
        _check_cuda_f32(n, t)
    if xa.dim() != 3 or xa.shape[1] != g.N or xa.shape[2] != g.Z:
        raise ValueError(f"xa must be [B, {g.N}, {g.Z}], got {tuple(xa.shape)}")
    if w.dim() != 2 or w.shape[1] != g.E or b.shape != w.shape:
        raise ValueError(f"w and b must be [T, {g.E}]")
    if xa.device.index != g.device_index:
        raise ValueError("graph handle and tensors live on different devices")
    return g, _aligned16(xa), w.contiguous(), b.contiguous()


def neural_forward_direct(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, graph_id: int) -> torch.Tensor:
    """Body of nldpc::neural_forward, callable without the dispatcher (inference under torch.no_grad())."""
    g, xa, w, b = _prep(xa, w, b, graph_id)
    B, T = xa.shape[0], w.shape[0]
    out = torch.empty((T, B, g.NZ), dtype=torch.float32, device=xa.device)
    with torch.cuda.device(xa.device):
        rc = _lib.lib().nldpc_neural_forward(g.ptr, _ptr(xa), _ptr(w), _ptr(b), B, T, _lib.NLDPC_OUT_ALL, _ptr(out),
                                             _lib.NLDPC_OUT_NONE, _vp(0), _stream(xa))
    _lib.check(rc, "nldpc_neural_forward")
    return out


@torch.library.custom_op("nldpc::neural_forward", mutates_args=())
def neural_forward(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, graph_id: int) -> torch.Tensor:
    return neural_forward_direct(xa, w, b, graph_id)


@neural_forward.register_fake
def _(xa, w, b, graph_id):
    g = _lib.graph_by_id(graph_id)
    return xa.new_empty((w.shape[0], xa.shape[0], g.NZ))


def neural_hard_direct(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, graph_id: int, all_iters: bool) -> torch.Tensor:
    """Body of nldpc::neural_hard, callable without the dispatcher (decode-only callers that hold no autograd state: the
    custom-op dispatch costs more than the whole batch-1024 launch)."""
    g, xa, w, b = _prep(xa, w, b, graph_id)
    B, T = xa.shape[0], w.shape[0]
    shape = (T, B, g.hard_bytes) if all_iters else (B, g.hard_bytes)
    hard = torch.empty(shape, dtype=torch.uint8, device=xa.device)
    with torch.cuda.device(xa.device):
        rc = _lib.lib().nldpc_neural_forward(g.ptr, _ptr(xa), _ptr(w), _ptr(b), B, T, _lib.NLDPC_OUT_NONE, _vp(0),
                                             _lib.NLDPC_OUT_ALL if all_iters else _lib.NLDPC_OUT_LAST, _ptr(hard), _stream(xa))
    _lib.check(rc, "nldpc_neural_forward")
    return hard


@torch.library.custom_op("nldpc::neural_hard", mutates_args=())
def neural_hard(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, graph_id: int, all_iters: bool) -> torch.Tensor:
    """Packed hard decisions (out < 0), uint8 [B, ceil(N*Z/8)] (last iteration) or [T, B, ...] (all_iters)."""
    return neural_hard_direct(xa, w, b, graph_id, all_iters)


@neural_hard.register_fake
def _(xa, w, b, graph_id, all_iters):
    g = _lib.graph_by_id(graph_id)
    shape = (w.shape[0], xa.shape[0], g.hard_bytes) if all_iters else (xa.shape[0], g.hard_bytes)
    return xa.new_empty(shape, dtype=torch.uint8)


@torch.library.custom_op("nldpc::neural_backward", mutates_args=())
def neural_backward(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, gout: torch.Tensor, graph_id: int,
                    dump: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """`dump`: the training dump written by nldpc::neural_forward_train for the same inputs (skips the forward re-run)"""
    g, xa, w, b = _prep(xa, w, b, graph_id)
    _check_cuda_f32("gout", gout)
    B, T = xa.shape[0], w.shape[0]
    if tuple(gout.shape) != (T, B, g.NZ):
        raise ValueError("gout must be [T, B, N*Z]")
    gout = gout.contiguous()
    gw = torch.empty_like(w)
    gb = torch.empty_like(b)
    nbytes = int(_lib.lib().nldpc_backward_workspace_bytes(g.ptr, B, T, 0))
    have_dump = dump is not None and dump.numel() >= nbytes
    ws = dump if have_dump else torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=xa.device)   # per-iteration v2c dump (HBM)
    with torch.cuda.device(xa.device):
        rc = _lib.lib().nldpc_neural_backward(g.ptr, _ptr(xa), _ptr(w), _ptr(b), _ptr(gout), B, T, _ptr(gw), _ptr(gb),
                                              _ptr(ws), nbytes, int(have_dump), _stream(xa))
    _lib.check(rc, "nldpc_neural_backward")
    return gw, gb


@neural_backward.register_fake
def _(xa, w, b, gout, graph_id, dump=None):
    return torch.empty_like(w), torch.empty_like(b)


@torch.library.custom_op("nldpc::neural_forward_train", mutates_args=())
def neural_forward_train(xa: torch.Tensor, w: torch.Tensor, b: torch.Tensor, graph_id: int) -> tuple[torch.Tensor, torch.Tensor]:
    """training-mode forward: out [T, B, N*Z] plus the per-iteration dump the backward kernel reads (uint8 workspace)"""
    g, xa, w, b = _prep(xa, w, b, graph_id)
    B, T = xa.shape[0], w.shape[0]
    out = torch.empty((T, B, g.NZ), dtype=torch.float32, device=xa.device)
    nbytes = int(_lib.lib().nldpc_backward_workspace_bytes(g.ptr, B, T, 0))
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=xa.device)
    with torch.cuda.device(xa.device):
        rc = _lib.lib().nldpc_neural_forward_train(g.ptr, _ptr(xa), _ptr(w), _ptr(b), B, T, _ptr(out), _ptr(ws), nbytes, _stream(xa))
    _lib.check(rc, "nldpc_neural_forward_train")
    return out, ws


@neural_forward_train.register_fake
def _(xa, w, b, graph_id):
    g = _lib.graph_by_id(graph_id)
    return xa.new_empty((w.shape[0], xa.shape[0], g.NZ)), xa.new_empty((1,), dtype=torch.uint8)


def _neural_train_setup_ctx(ctx, inputs, output):
    xa, w, b, graph_id = inputs
    ctx.save_for_backward(xa, w, b, output[1])
    ctx.graph_id = graph_id
    ctx.set_materialize_grads(False)     # no zero tensor the size of the training dump for its (unused) gradient


def _neural_train_bwd(ctx, gout, gws):
    if gout is None:
        return None, None, None, None
    xa, w, b, ws = ctx.saved_tensors
    gw, gb = torch.ops.nldpc.neural_backward(xa, w, b, gout.contiguous(), ctx.graph_id, ws)
    return None, gw, gb, None


neural_forward_train.register_autograd(_neural_train_bwd, setup_context=_neural_train_setup_ctx)


def _neural_setup_ctx(ctx, inputs, output):
    xa, w, b, graph_id = inputs
    ctx.save_for_backward(xa, w, b)
    ctx.graph_id = graph_id


def _neural_bwd(ctx, gout):
    xa, w, b = ctx.saved_tensors
    gw, gb = torch.ops.nldpc.neural_backward(xa, w, b, gout.contiguous(), ctx.graph_id, None)
    return None, gw, gb, None


neural_forward.register_autograd(_neural_bwd, setup_context=_neural_setup_ctx)


def neural_decode_host(graph_id, xa_host, w_host, b_host, soft_mode=_lib.NLDPC_OUT_NONE, hard_mode=_lib.NLDPC_OUT_LAST, scale=1.0):
    """End-to-end host-buffer decode (nldpc_neural_decode_host): CPU tensors in, CPU tensors out; the
    H2D copy, the kernel and the D2H copy of consecutive chunks overlap inside the library.
    xa_host float32, or — the path is host->device-link bound — float16 values / int8 codes (x = scale * q), expanded on the
    device (nldpc_neural_decode_host_narrow): bit-identical to decoding the widened values."""
    g = _lib.graph_by_id(graph_id)
    if xa_host.is_cuda or xa_host.dtype not in (torch.float32, torch.float16, torch.int8) or not xa_host.is_contiguous():
        raise ValueError("xa must be a contiguous float32 / float16 / int8 CPU tensor")
    if tuple(xa_host.shape[1:]) != (g.N, g.Z):
        raise ValueError(f"xa must be [B, {g.N}, {g.Z}], got {tuple(xa_host.shape)}")
    for n, t in (("w", w_host), ("b", b_host)):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{n} must be a contiguous float32 CPU tensor")
    B, T = xa_host.shape[0], w_host.shape[0]
    soft = hard = None
    if soft_mode != _lib.NLDPC_OUT_NONE:
        soft = torch.empty((T, B, g.NZ) if soft_mode == _lib.NLDPC_OUT_ALL else (B, g.NZ), dtype=torch.float32,
                           pin_memory=True)
    if hard_mode != _lib.NLDPC_OUT_NONE:
        hard = torch.empty((T, B, g.hard_bytes) if hard_mode == _lib.NLDPC_OUT_ALL else (B, g.hard_bytes),
                           dtype=torch.uint8, pin_memory=True)
    if xa_host.dtype == torch.float32:
        rc = _lib.lib().nldpc_neural_decode_host(g.ptr, _ptr(xa_host), _ptr(w_host), _ptr(b_host), B, T, soft_mode, _ptr(soft),
                                                 hard_mode, _ptr(hard))
    else:
        fmt = _lib.NLDPC_LLR_F16 if xa_host.dtype == torch.float16 else _lib.NLDPC_LLR_Q8
        rc = _lib.lib().nldpc_neural_decode_host_narrow(g.ptr, _ptr(xa_host), fmt, float(scale), _ptr(w_host), _ptr(b_host), B, T,
                                                        soft_mode, _ptr(soft), hard_mode, _ptr(hard))
    _lib.check(rc, "nldpc_neural_decode_host")
    return soft, hard


# ---------------------------------------------------------------------------------------------------------------
# Boosted decoder (BoostedNeuralLDPCDecoder.py:320-531)


def _out_shape(mode, T, B, last):
    return (T, B, last) if mode == _lib.NLDPC_OUT_ALL else ((B, last) if mode == _lib.NLDPC_OUT_LAST else (0,))


def _llr_shape(llr_mode, T, B, g):
    return (T, B, g.Z, g.E) if llr_mode == 2 else ((B, g.Z, g.E) if llr_mode == 1 else (0,))


def _opt_f32(name, t, shape, device):
    if t is None:
        return None
    _check_cuda_f32(name, t)
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, expected {device}")
    return t.contiguous()


def boosted_forward_direct(xa, vn_w, cn_w, ucn_w, graph_id, T, decoder_type, qbit, llr_lo, llr_hi, compute_ucn, ucn_mix, llr_init,
                           xin_init, app_init, llr_mode, want_xin, soft_mode, hard_mode, want_dump=False, pad_llr=False):
    """Body of nldpc::boosted_forward, callable without the dispatcher (decode-only callers that hold no autograd state).
    pad_llr: True = allocate the llr state with 16-byte rows (row pitch E rounded up to a multiple of 4 floats, nldpc_boosted_cfg_t
    llr_pitch), an int = that row pitch; returns the [.., :E] view of it (same shape and values; not contiguous when pitch != E).  16-byte rows are what
    the vector state export of the specialised kernels needs (graphs without degree-1 blocks; WiMAX's E = 88 has them as is).
    T consecutive iterations of the Boosted loop body.  Returns (soft [T,B,N*Z] | [B,N*Z] | empty per soft_mode,
    llr [B,Z,E] (llr_mode 1: self.llr[t_last+1]) | [T,B,Z,E] (llr_mode 2: every executed iteration, :512) | empty (0),
    xin_out [B,N,Z] | empty, packed hard decisions per hard_mode | empty, training dump | 1 byte).
    Weight rows are indexed by executed iteration."""
    g = _lib.graph_by_id(graph_id)
    _check_cuda_f32("xa", xa)
    if xa.dim() != 3 or xa.shape[1] != g.N or xa.shape[2] != g.Z:
        raise ValueError(f"xa must be [B, {g.N}, {g.Z}], got {tuple(xa.shape)}")
    if xa.device.index != g.device_index:
        raise ValueError("graph handle and tensors live on different devices")
    B, dev = xa.shape[0], xa.device
    xa = _aligned16(xa)
    vn_w = _opt_f32("vn_w", vn_w, (T, g.N), dev)
    cn_w = _opt_f32("cn_w", cn_w, (T, g.E), dev)
    ucn_w = _opt_f32("ucn_w", ucn_w, (T, g.E), dev)
    llr_init = _opt_f32("llr_init", llr_init, (B, g.Z, g.E), dev)
    xin_init = _opt_f32("xin_init", xin_init, (B, g.N, g.Z), dev)
    app_init = _opt_f32("app_init", app_init, (B, g.NZ), dev)
    soft = torch.empty(_out_shape(soft_mode, T, B, g.NZ), dtype=torch.float32, device=dev)
    hard = torch.empty(_out_shape(hard_mode, T, B, g.hard_bytes), dtype=torch.uint8, device=dev)
    llr_mode = int(llr_mode)
    if pad_llr is True:
        llr_pitch = ((g.E + 3) // 4) * 4 if llr_mode else g.E
    else:
        llr_pitch = int(pad_llr) if (pad_llr and llr_mode) else g.E      # an explicit row pitch in floats (the library rejects < E)
    llr_shape = _llr_shape(llr_mode, T, B, g)
    llr_last = torch.empty(llr_shape[:-1] + (max(llr_pitch, 1),) if llr_mode else llr_shape, dtype=torch.float32, device=dev)
    xin_out = torch.empty((B, g.N, g.Z) if want_xin else (0,), dtype=torch.float32, device=dev)
    nbytes = int(_lib.lib().nldpc_backward_workspace_bytes(g.ptr, B, T, 1)) if want_dump else 0
    dump = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    cfg = _lib.BoostedCfg(decoder_type, qbit, llr_lo, llr_hi, int(compute_ucn), int(ucn_mix),
                          llr_init.data_ptr() if llr_init is not None else None,
                          xin_init.data_ptr() if xin_init is not None else None,
                          xin_out.data_ptr() if want_xin else None,
                          app_init.data_ptr() if app_init is not None else None,
                          dump.data_ptr() if want_dump else None, nbytes,
                          llr_last.data_ptr() if llr_mode == 2 else None, llr_pitch)
    with torch.cuda.device(dev):
        rc = _lib.lib().nldpc_boosted_forward(g.ptr, ctypes.byref(cfg), _ptr(xa), _ptr(vn_w), _ptr(cn_w), _ptr(ucn_w), B, T,
                                              soft_mode, _ptr(soft) if soft_mode else _vp(0), hard_mode,
                                              _ptr(hard) if hard_mode else _vp(0),
                                              _ptr(llr_last) if llr_mode == 1 else _vp(0), _stream(xa))
    _lib.check(rc, "nldpc_boosted_forward")
    if llr_mode and llr_pitch != g.E:
        llr_last = llr_last[..., :g.E]
    return soft, llr_last, xin_out, hard, dump


@torch.library.custom_op("nldpc::boosted_forward", mutates_args=())
def boosted_forward(xa: torch.Tensor, vn_w: Optional[torch.Tensor], cn_w: Optional[torch.Tensor], ucn_w: Optional[torch.Tensor],
                    graph_id: int, T: int, decoder_type: int, qbit: int, llr_lo: float, llr_hi: float, compute_ucn: bool,
                    ucn_mix: bool, llr_init: Optional[torch.Tensor], xin_init: Optional[torch.Tensor],
                    app_init: Optional[torch.Tensor], llr_mode: int, want_xin: bool, soft_mode: int,
                    hard_mode: int, want_dump: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """see boosted_forward_direct"""
    return boosted_forward_direct(xa, vn_w, cn_w, ucn_w, graph_id, T, decoder_type, qbit, llr_lo, llr_hi, compute_ucn, ucn_mix, llr_init,
                                  xin_init, app_init, llr_mode, want_xin, soft_mode, hard_mode, want_dump)


@boosted_forward.register_fake
def _(xa, vn_w, cn_w, ucn_w, graph_id, T, decoder_type, qbit, llr_lo, llr_hi, compute_ucn, ucn_mix, llr_init, xin_init, app_init,
      llr_mode, want_xin, soft_mode, hard_mode, want_dump=False):
    g = _lib.graph_by_id(graph_id)
    B = xa.shape[0]
    return (xa.new_empty(_out_shape(soft_mode, T, B, g.NZ)), xa.new_empty(_llr_shape(llr_mode, T, B, g)),
            xa.new_empty((B, g.N, g.Z) if want_xin else (0,)),
            xa.new_empty(_out_shape(hard_mode, T, B, g.hard_bytes), dtype=torch.uint8), xa.new_empty((1,), dtype=torch.uint8))


@torch.library.custom_op("nldpc::boosted_backward", mutates_args=())
def boosted_backward(xa: torch.Tensor, vn_w: Optional[torch.Tensor], cn_w: Optional[torch.Tensor], ucn_w: Optional[torch.Tensor],
                     gout: torch.Tensor, graph_id: int, T: int, decoder_type: int, qbit: int, llr_lo: float, llr_hi: float,
                     compute_ucn: bool, ucn_mix: bool, dump: Optional[torch.Tensor] = None,
                     dump_fmt: int = 0) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """gradients w.r.t. the folded weight rows (vn [T,N], cn [T,E], ucn [T,E]; empty where absent).  `dump`: the training dump
    the forward op wrote (want_dump), `dump_fmt` its format (nldpc_boosted_dump_format of that forward call)."""
    g = _lib.graph_by_id(graph_id)
    _check_cuda_f32("xa", xa)
    _check_cuda_f32("gout", gout)
    B, dev = xa.shape[0], xa.device
    if tuple(gout.shape) != (T, B, g.NZ):
        raise ValueError("gout must be [T, B, N*Z]")
    xa, gout = _aligned16(xa), gout.contiguous()
    vn_w = _opt_f32("vn_w", vn_w, (T, g.N), dev)
    cn_w = _opt_f32("cn_w", cn_w, (T, g.E), dev)
    ucn_w = _opt_f32("ucn_w", ucn_w, (T, g.E), dev)
    gvn = torch.empty((T, g.N) if vn_w is not None else (0,), dtype=torch.float32, device=dev)
    gcn = torch.empty((T, g.E) if cn_w is not None else (0,), dtype=torch.float32, device=dev)
    gucn = torch.empty((T, g.E) if (ucn_w is not None and ucn_mix) else (0,), dtype=torch.float32, device=dev)
    nbytes = int(_lib.lib().nldpc_backward_workspace_bytes(g.ptr, B, T, 1))
    have_dump = dump is not None and dump.numel() >= nbytes
    ws = dump if have_dump else torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=dev)
    cfg = _lib.BoostedCfg(decoder_type, qbit, llr_lo, llr_hi, int(compute_ucn), int(ucn_mix), None, None, None, None, None, 0, None)
    with torch.cuda.device(dev):
        rc = _lib.lib().nldpc_boosted_backward(g.ptr, ctypes.byref(cfg), _ptr(xa), _ptr(vn_w), _ptr(cn_w), _ptr(ucn_w), _ptr(gout), B, T,
                                               _ptr(gvn) if gvn.numel() else _vp(0), _ptr(gcn) if gcn.numel() else _vp(0),
                                               _ptr(gucn) if gucn.numel() else _vp(0), _ptr(ws), nbytes,
                                               (1 + int(dump_fmt)) if have_dump else 0, _stream(xa))
    _lib.check(rc, "nldpc_boosted_backward")
    return gvn, gcn, gucn


@boosted_backward.register_fake
def _(xa, vn_w, cn_w, ucn_w, gout, graph_id, T, decoder_type, qbit, llr_lo, llr_hi, compute_ucn, ucn_mix, dump=None, dump_fmt=0):
    g = _lib.graph_by_id(graph_id)
    return (xa.new_empty((T, g.N) if vn_w is not None else (0,)), xa.new_empty((T, g.E) if cn_w is not None else (0,)),
            xa.new_empty((T, g.E) if (ucn_w is not None and ucn_mix) else (0,)))


def _boosted_setup_ctx(ctx, inputs, output):
    (xa, vn_w, cn_w, ucn_w, graph_id, T, dec, qbit, lo, hi, compute_ucn, ucn_mix, llr_init, xin_init, app_init, llr_mode,
     want_xin, soft_mode, hard_mode, want_dump) = inputs
    ctx.save_for_backward(xa, vn_w, cn_w, ucn_w, output[4] if want_dump else None)
    ctx.cfg = (graph_id, T, dec, qbit, lo, hi, compute_ucn, ucn_mix)
    ctx.stateful = llr_init is not None or xin_init is not None or app_init is not None
    ctx.have_dump = bool(want_dump)
    ctx.dump_fmt = 0
    if want_dump:       # the format the forward call wrote (its state pointers matter: stateful runs use the table-driven kernels)
        g = _lib.graph_by_id(graph_id)
        one = ctypes.c_void_p(1)
        cfg = _lib.BoostedCfg(dec, qbit, lo, hi, int(compute_ucn), int(ucn_mix), one if llr_init is not None else None,
                              one if xin_init is not None else None, one if want_xin else None,
                              one if app_init is not None else None, None, 0, None)
        ctx.dump_fmt = int(_lib.lib().nldpc_boosted_dump_format(g.ptr, ctypes.byref(cfg), T, int(cn_w is not None), int(vn_w is not None)))
    ctx.soft_all = soft_mode == _lib.NLDPC_OUT_ALL
    # autograd would otherwise hand the backward a ZERO tensor for every output without a gradient — including one the size
    # of the training dump (20 GB at B = 65536, T = 20: 5 ms of fill per step)
    ctx.set_materialize_grads(False)


def _boosted_bwd(ctx, gsoft, gllr, gxin, ghard, gdump):
    if gsoft is None:
        return (None,) * 20
    xa, vn_w, cn_w, ucn_w, dump = ctx.saved_tensors
    graph_id, T, dec, qbit, lo, hi, compute_ucn, ucn_mix = ctx.cfg
    if not ctx.soft_all:
        raise _lib.NldpcError("backward needs the per-iteration soft outputs (soft_mode = ALL)")
    if ctx.stateful and not ctx.have_dump:
        # the stored state (llr_init / xin_init / app_init) is a constant of the call, exactly as the graph-less tensors the
        # reference reads from self.llr / self.outputs are; the sweep only needs the dump the forward wrote WITH that state
        raise _lib.NldpcError("backward through a run that continues from stored state needs the training dump of that run "
                              "(call the forward op with want_dump=True)")
    gvn, gcn, gucn = torch.ops.nldpc.boosted_backward(xa, vn_w, cn_w, ucn_w, gsoft.contiguous(), graph_id, T, dec, qbit, lo, hi,
                                                      compute_ucn, ucn_mix, dump, ctx.dump_fmt)
    return (None, gvn if vn_w is not None else None, gcn if cn_w is not None else None,
            gucn if (ucn_w is not None and ucn_mix) else None) + (None,) * 16


boosted_forward.register_autograd(_boosted_bwd, setup_context=_boosted_setup_ctx)


# ---------------------------------------------------------------------------------------------------------------
# fused training step: forward + multi-iteration BCE + dL/dout in one launch, the sweep reads the workspace
def pack_labels(y: torch.Tensor) -> torch.Tensor:
    """y [B, N*Z] fp32 labels (0 / 1) -> uint8 [B, ceil(N*Z/8)], bit i of a codeword = (y[i] != 0), LSB first — the packing of
    the decode ops' hard decisions."""
    _check_cuda_f32("y", y)
    if y.dim() != 2:
        raise ValueError("y must be [B, N*Z]")
    y = y.contiguous()
    bits = torch.empty((y.shape[0], (y.shape[1] + 7) // 8), dtype=torch.uint8, device=y.device)
    with torch.cuda.device(y.device):
        rc = _lib.lib().nldpc_pack_labels(_ptr(y), y.shape[0], y.shape[1], _ptr(bits), _stream(y))
    _lib.check(rc, "nldpc_pack_labels")
    return bits


def _train_cfg(decoder_type, qbit, llr_lo, llr_hi):
    return _lib.BoostedCfg(decoder_type, qbit, llr_lo, llr_hi, 0, 0, None, None, None, None, None, 0, None)


def boosted_train_covered(graph_id, decoder_type, qbit, llr_lo, llr_hi, T, has_cn_w, has_vn_w) -> bool:
    """does the fused training path (boosted_train_loss) take this configuration?"""
    g = _lib.graph_by_id(graph_id)
    cfg = _train_cfg(decoder_type, qbit, llr_lo, llr_hi)
    return int(_lib.lib().nldpc_boosted_train_workspace_bytes(g.ptr, ctypes.byref(cfg), 1, T, int(has_cn_w), int(has_vn_w))) > 0


class _BoostedTrainLoss(torch.autograd.Function):
    """L = sum_t coef[t] * mean bce_with_logits(out_t, y) of the Boosted decoder run from the zero state, differentiable w.r.t.
    the folded weight rows.  forward: nldpc_boosted_train_forward (one launch: T iterations + loss + dL/dout + dump);
    backward: nldpc_boosted_train_backward (the sweep)."""

    @staticmethod
    def forward(ctx, xa, vn_w, cn_w, ybits, coef, graph_id, T, decoder_type, qbit, llr_lo, llr_hi):
        g = _lib.graph_by_id(graph_id)
        _check_cuda_f32("xa", xa)
        if xa.dim() != 3 or xa.shape[1] != g.N or xa.shape[2] != g.Z:
            raise ValueError(f"xa must be [B, {g.N}, {g.Z}], got {tuple(xa.shape)}")
        B, dev = xa.shape[0], xa.device
        if dev.index != g.device_index:
            raise ValueError("graph handle and tensors live on different devices")
        xa = _aligned16(xa)
        vn_w = _opt_f32("vn_w", vn_w, (T, g.N), dev)
        cn_w = _opt_f32("cn_w", cn_w, (T, g.E), dev)
        _check_cuda_f32("coef", coef)
        if ybits.dtype != torch.uint8 or tuple(ybits.shape) != (B, g.hard_bytes) or ybits.device != dev or coef.numel() != T:
            raise ValueError("ybits must be uint8 [B, ceil(N*Z/8)] (ops.pack_labels) and coef [T]")
        ybits, coef = ybits.contiguous(), coef.contiguous()
        cfg = _train_cfg(decoder_type, qbit, llr_lo, llr_hi)
        nbytes = int(_lib.lib().nldpc_boosted_train_workspace_bytes(g.ptr, ctypes.byref(cfg), B, T, int(cn_w is not None), int(vn_w is not None)))
        if nbytes == 0:
            raise _lib.NldpcError("boosted_train_loss: configuration not covered by the fused training kernels (boosted_train_covered)")
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        loss_sum = torch.empty((1,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().nldpc_boosted_train_forward(g.ptr, ctypes.byref(cfg), _ptr(xa), _ptr(vn_w), _ptr(cn_w), B, T, _ptr(ybits),
                                                        _ptr(coef), 1.0, _ptr(loss_sum), _ptr(ws), nbytes, _stream(xa))
        _lib.check(rc, "nldpc_boosted_train_forward")
        ctx.save_for_backward(xa, vn_w, cn_w, ws)
        ctx.cfg = (graph_id, T, decoder_type, qbit, llr_lo, llr_hi)
        ctx.set_materialize_grads(False)
        return (loss_sum[0] / float(B * g.NZ)).to(torch.float32)

    @staticmethod
    def backward(ctx, gloss):
        if gloss is None:
            return (None,) * 11
        xa, vn_w, cn_w, ws = ctx.saved_tensors
        graph_id, T, decoder_type, qbit, llr_lo, llr_hi = ctx.cfg
        g = _lib.graph_by_id(graph_id)
        B, dev = xa.shape[0], xa.device
        cfg = _train_cfg(decoder_type, qbit, llr_lo, llr_hi)
        gvn = torch.empty((T, g.N), dtype=torch.float32, device=dev) if vn_w is not None else None
        gcn = torch.empty((T, g.E), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.lib().nldpc_boosted_train_backward(g.ptr, ctypes.byref(cfg), _ptr(xa), _ptr(vn_w), _ptr(cn_w), B, T, _ptr(gvn),
                                                         _ptr(gcn), _ptr(ws), ws.numel(), _stream(xa))
        _lib.check(rc, "nldpc_boosted_train_backward")
        gl = gloss.to(torch.float32)            # the sweep ran with upstream 1; the weight gradients are linear in it
        return (None, gvn * gl if gvn is not None else None, gcn * gl) + (None,) * 8


def boosted_train_loss(xa, vn_w, cn_w, ybits, coef, graph_id, T, decoder_type, qbit, llr_lo, llr_hi):
    """see _BoostedTrainLoss; vn_w [T,N] | None, cn_w [T,E] folded rows, ybits from pack_labels, coef [T] normalised weights"""
    return _BoostedTrainLoss.apply(xa, vn_w, cn_w, ybits, coef, graph_id, T, decoder_type, qbit, llr_lo, llr_hi)


def iteration_coefs(n, etha, coeff_param, device):
    """normalised iteration weights etha^{c_t} / sum_t etha^{c_t} of LDPCDecoderLoss (LDPCDecoderLoss.py:88-106) as a cached
    device tensor (one H2D copy per distinct vector, none per step: CUDA-graph capturable)"""
    if coeff_param is None:
        coeffs = [1] * n
    else:
        coeffs = list(coeff_param) if isinstance(coeff_param, (list, tuple)) else [coeff_param] * n
    w = [float(pow(etha, c)) for c in coeffs]
    tot = sum(w)
    key = (device, tuple(v / tot if tot > 0 else v for v in w))
    coef = _coef_cache.get(key)
    if coef is None:
        if len(_coef_cache) > 64:
            _coef_cache.clear()
        coef = _coef_cache[key] = torch.tensor(key[1], dtype=torch.float32, device=device)
    return coef


_coef_cache = {}


# ---------------------------------------------------------------------------------------------------------------
# fused multi-iteration BCE (LDPCDecoderLoss.py:73-108, BCE branch)
@torch.library.custom_op("nldpc::multi_iter_bce", mutates_args=())
def multi_iter_bce(soft: torch.Tensor, y: torch.Tensor, coef: torch.Tensor, want_grad: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """soft [T, ...] logits, y [...] labels, coef [T] normalised iteration weights -> (loss scalar, dL/dsoft or empty)"""
    _check_cuda_f32("soft", soft)
    _check_cuda_f32("y", y)
    _check_cuda_f32("coef", coef)
    T = soft.shape[0]
    n = y.numel()
    if soft.numel() != T * n or coef.numel() != T:
        raise ValueError("soft must be [T, *y.shape] and coef [T]")
    soft, y, coef = soft.contiguous(), y.contiguous(), coef.contiguous()
    loss = torch.empty((), dtype=torch.float32, device=soft.device)
    gout = torch.empty_like(soft) if want_grad else torch.empty((0,), dtype=torch.float32, device=soft.device)
    with torch.cuda.device(soft.device):
        rc = _lib.lib().nldpc_multi_iter_bce(_ptr(soft), _ptr(y), _ptr(coef), T, n, _ptr(loss), _ptr(gout) if want_grad else _vp(0),
                                             _stream(soft))
    _lib.check(rc, "nldpc_multi_iter_bce")
    return loss, gout


@multi_iter_bce.register_fake
def _(soft, y, coef, want_grad):
    return soft.new_empty(()), (torch.empty_like(soft) if want_grad else soft.new_empty((0,)))


@torch.library.custom_op("nldpc::multi_iter_bce_grad", mutates_args=())
def multi_iter_bce_grad(soft: torch.Tensor, y: torch.Tensor, coef: torch.Tensor, gscale: torch.Tensor) -> torch.Tensor:
    """dL/dsoft of multi_iter_bce with the upstream scalar gradient `gscale` (a device tensor) folded in: one pass, no
    [T, ...] tensor kept between forward and backward"""
    _check_cuda_f32("soft", soft)
    _check_cuda_f32("y", y)
    _check_cuda_f32("coef", coef)
    _check_cuda_f32("gscale", gscale)
    T, n = soft.shape[0], y.numel()
    if soft.numel() != T * n or coef.numel() != T or gscale.numel() != 1:
        raise ValueError("soft must be [T, *y.shape], coef [T], gscale a scalar")
    soft, y, coef, gscale = soft.contiguous(), y.contiguous(), coef.contiguous(), gscale.contiguous()
    gout = torch.empty_like(soft)
    with torch.cuda.device(soft.device):
        rc = _lib.lib().nldpc_multi_iter_bce_grad(_ptr(soft), _ptr(y), _ptr(coef), _ptr(gscale), T, n, _ptr(gout), _stream(soft))
    _lib.check(rc, "nldpc_multi_iter_bce_grad")
    return gout


@multi_iter_bce_grad.register_fake
def _(soft, y, coef, gscale):
    return torch.empty_like(soft)


def _bce_setup_ctx(ctx, inputs, output):
    soft, y, coef, want_grad = inputs
    ctx.save_for_backward(soft, y, coef)
    ctx.set_materialize_grads(False)


def _bce_bwd(ctx, gloss, ggout):
    if gloss is None:
        return None, None, None, None
    soft, y, coef = ctx.saved_tensors
    return torch.ops.nldpc.multi_iter_bce_grad(soft, y, coef, gloss.to(torch.float32)), None, None, None


multi_iter_bce.register_autograd(_bce_bwd, setup_context=_bce_setup_ctx)


def fused_multi_iter_bce(outputs, y, etha=1.0, coeff_param=None):
    """Drop-in for LDPCDecoderLoss(BCE)(outputs, y, coeff_param) when `outputs` is the list a decoder's forward returned
    (views of ONE [T, B, N*Z] tensor): a single fused kernel for the loss and its gradient.  Returns None when the list
    is not such a set of views (caller falls back to the per-iteration torch ops)."""
    if not isinstance(outputs, (list, tuple)) or len(outputs) == 0 or not all(isinstance(o, torch.Tensor) and o.is_cuda for o in outputs):
        return None
    base = outputs[0]._base
    if base is None or base.dim() != outputs[0].dim() + 1 or base.shape[0] != len(outputs) or not base.is_contiguous():
        return None
    step = base.stride(0)
    for t, o in enumerate(outputs):
        if o._base is not base or o.storage_offset() != base.storage_offset() + t * step or tuple(o.shape) != tuple(base.shape[1:]):
            return None
    coef = iteration_coefs(len(outputs), etha, coeff_param, base.device)
    loss, _ = torch.ops.nldpc.multi_iter_bce(base, y.to(torch.float32), coef, False)    # the gradient is produced in backward
    return 1.0 * loss


# ---------------------------------------------------------------------------------------------------------------
# bit / frame error counts per iteration (Functions.evaluate_ber_fer, Functions.py:86-102)
@torch.library.custom_op("nldpc::count_errors", mutates_args=())
def count_errors(soft: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """soft [T, B, NZ] iteration outputs, y [B, NZ] fp32 labels -> int64 [2, T]: row 0 = positions where
    ((out < 0) ? 1 : 0) != y, row 1 = codewords with at least one (one fused pass; exact integer counts)."""
    _check_cuda_f32("soft", soft)
    _check_cuda_f32("y", y)
    if soft.dim() != 3 or y.dim() != 2 or tuple(soft.shape[1:]) != tuple(y.shape):
        raise ValueError("soft must be [T, B, NZ] and y [B, NZ]")
    T, B, NZ = soft.shape
    if T == 0:
        return torch.zeros((2, 0), dtype=torch.int64, device=soft.device)
    if not (soft.stride(2) == 1 and soft.stride(1) == NZ) and B * NZ > 0:      # rows must be dense; the iteration stride is free
        soft = soft.contiguous()
    y = y.contiguous()
    counts = torch.empty((2, T), dtype=torch.int64, device=soft.device)
    with torch.cuda.device(soft.device):
        rc = _lib.lib().nldpc_count_errors(_ptr(soft), soft.stride(0) if B * NZ > 0 else 0, _ptr(y), T, B, NZ, _ptr(counts), _stream(soft))
    _lib.check(rc, "nldpc_count_errors")
    return counts


@count_errors.register_fake
def _(soft, y):
    return soft.new_empty((2, soft.shape[0]), dtype=torch.int64)


@torch.library.custom_op("nldpc::count_errors_packed", mutates_args=())
def count_errors_packed(hard: torch.Tensor, n_bits: int, y_packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """hard uint8 [T, B, ceil(n_bits/8)] (or [B, ...] = one iteration) packed decisions as the decode ops write them,
    y_packed the labels packed the same way (None = all-zero codeword) -> int64 [2, T] as count_errors."""
    if not hard.is_cuda:
        raise _lib.NldpcError(f"hard is on {hard.device}: the B200 path has no CPU fallback — move it to a CUDA device")
    if hard.dtype != torch.uint8 or (y_packed is not None and y_packed.dtype != torch.uint8):
        raise TypeError("packed decisions / labels must be uint8")
    if hard.dim() == 2:
        hard = hard.unsqueeze(0)
    hb = (n_bits + 7) // 8
    if hard.dim() != 3 or hard.shape[2] != hb or (y_packed is not None and tuple(y_packed.shape) != tuple(hard.shape[1:])):
        raise ValueError("hard must be [T, B, ceil(n_bits/8)] and y_packed [B, ceil(n_bits/8)]")
    T, B, _ = hard.shape
    if T == 0:
        return torch.zeros((2, 0), dtype=torch.int64, device=hard.device)
    hard = hard.contiguous()
    yp = y_packed.contiguous() if y_packed is not None else None
    counts = torch.empty((2, T), dtype=torch.int64, device=hard.device)
    with torch.cuda.device(hard.device):
        rc = _lib.lib().nldpc_count_errors_packed(_ptr(hard), B * hb, _ptr(yp), T, B, n_bits, _ptr(counts), _stream(hard))
    _lib.check(rc, "nldpc_count_errors_packed")
    return counts


@count_errors_packed.register_fake
def _(hard, n_bits, y_packed=None):
    return hard.new_empty((2, hard.shape[0] if hard.dim() == 3 else 1), dtype=torch.int64)


def fused_ber_fer_counts(expected, actual):
    """Functions.evaluate_ber_fer on the device: `actual` is a list of T CUDA fp32 [B, NZ] tensors (views of one [T, B, NZ]
    tensor are read in place; anything else is counted tensor by tensor, still one pass each).  Returns int64 [2, T] on the
    device, or None when the inputs are not CUDA fp32 2-D tensors of one shape (caller uses the torch formulation)."""
    if not isinstance(actual, (list, tuple)) or len(actual) == 0 or not isinstance(expected, torch.Tensor):
        return None
    if not (expected.is_cuda and expected.dtype == torch.float32 and expected.dim() == 2):
        return None
    for o in actual:
        if not (isinstance(o, torch.Tensor) and o.is_cuda and o.dtype == torch.float32 and tuple(o.shape) == tuple(expected.shape)
                and o.device == expected.device):
            return None
    base = actual[0]._base
    if base is not None and base.dim() == 3 and base.shape[0] == len(actual) and tuple(base.shape[1:]) == tuple(expected.shape):
        step = base.stride(0)
        if all(o._base is base and o.storage_offset() == base.storage_offset() + t * step and o.stride() == base.stride()[1:]
               for t, o in enumerate(actual)):
            return torch.ops.nldpc.count_errors(base.detach(), expected)
    return torch.cat([torch.ops.nldpc.count_errors(o.detach().unsqueeze(0), expected) for o in actual], dim=1)


# ---------------------------------------------------------------------------------------------------------------
# clip_grad_norm_ + Adam + clamp on the flat weight vector (train/train_BoostedNeuralLDPCDecoder.py:291-294)
def clip_adam_clamp_(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, state: torch.Tensor,
                     grad_scale: float = 1.0, max_norm: float = 1.0, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                     clamp=(0.0, 2.0), lr_dev: Optional[torch.Tensor] = None):
    """In place, one launch, no host synchronisation (CUDA-graph replayable: the step counter is state[0] on the device, the
    learning rate is read from `lr_dev` when given).  `grad` may be longer than `param`: the extra entries (gradients of
    parameters the optimiser does not update) enter the clipping norm and are scaled, as clip_grad_norm_(model.parameters())
    does in the reference (train/train_BoostedNeuralLDPCDecoder.py:291)."""
    n = param.numel()
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq), ("state", state)):
        _check_cuda_f32(name, t)
        if not t.is_contiguous() or (name not in ("state", "grad") and t.numel() != n):
            raise ValueError(f"{name} must be a contiguous fp32 vector of {n} elements")
    if grad.numel() < n:
        raise ValueError("grad must hold at least as many elements as param")
    if state.numel() < 2:
        raise ValueError("state must hold at least 2 floats (step count, last gradient norm)")
    if lr_dev is not None:
        _check_cuda_f32("lr_dev", lr_dev)
    with torch.cuda.device(param.device):
        rc = _lib.lib().nldpc_clip_adam_clamp(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(state), n, grad.numel(),
                                              float(grad_scale), float(max_norm), float(lr), _ptr(lr_dev), float(betas[0]),
                                              float(betas[1]), float(eps), float(clamp[0]), float(clamp[1]), _stream(param))
    _lib.check(rc, "nldpc_clip_adam_clamp")


# ---------------------------------------------------------------------------------------------------------------
# host-buffer Boosted decode with one-byte channel LLRs (nldpc_boosted_decode_host_q8)
def boosted_decode_host_q8(graph_id: int, xq_host: torch.Tensor, scale: float, vn_w_host, cn_w_host, ucn_w_host, T: int,
                           decoder_type: int, qbit: int, llr_lo: float, llr_hi: float, compute_ucn: bool, ucn_mix: bool,
                           soft_mode: int = 0, hard_mode: int = 2):
    """xq_host int8 CPU tensor [B, N, Z] (x = scale * q; pin it for full PCIe speed), folded weight rows as CPU fp32 tensors
    or None -> (soft | None, hard | None) as pinned CPU tensors, shapes as the device op.  Synchronous."""
    g = _lib.graph_by_id(graph_id)
    if not isinstance(xq_host, torch.Tensor) or xq_host.is_cuda or xq_host.dtype != torch.int8:
        raise TypeError("xq_host must be an int8 CPU tensor")
    if xq_host.dim() != 3 or xq_host.shape[1] != g.N or xq_host.shape[2] != g.Z:
        raise ValueError(f"xq_host must be [B, {g.N}, {g.Z}], got {tuple(xq_host.shape)}")
    xq_host = xq_host.contiguous()
    B = xq_host.shape[0]

    def host_rows(name, t, cols):
        if t is None:
            return None
        if t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (T, cols):
            raise ValueError(f"{name} must be a CPU fp32 tensor [{T}, {cols}]")
        return t.contiguous()

    vn_w_host, cn_w_host, ucn_w_host = host_rows("vn_w", vn_w_host, g.N), host_rows("cn_w", cn_w_host, g.E), host_rows("ucn_w", ucn_w_host, g.E)
    soft = hard = None
    if soft_mode != _lib.NLDPC_OUT_NONE:
        soft = torch.empty(_out_shape(soft_mode, T, B, g.NZ), dtype=torch.float32, pin_memory=True)
    if hard_mode != _lib.NLDPC_OUT_NONE:
        hard = torch.empty(_out_shape(hard_mode, T, B, g.hard_bytes), dtype=torch.uint8, pin_memory=True)
    cfg = _lib.BoostedCfg(decoder_type, qbit, llr_lo, llr_hi, int(compute_ucn), int(ucn_mix), None, None, None, None, None, 0, None)
    rc = _lib.lib().nldpc_boosted_decode_host_q8(g.ptr, ctypes.byref(cfg), _ptr(xq_host), float(scale), _ptr(vn_w_host), _ptr(cn_w_host),
                                                 _ptr(ucn_w_host), B, T, soft_mode, _ptr(soft), hard_mode, _ptr(hard))
    _lib.check(rc, "nldpc_boosted_decode_host_q8")
    return soft, hard
