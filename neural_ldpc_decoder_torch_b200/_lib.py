"""ctypes loader of libnldpc_b200.so (the C ABI of include/nldpc.h) and per-device graph handles.

There is NO fallback: if the library is missing or no sm_100 device is present the calls raise.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NLDPC_LIB_PATH") or os.path.join(_HERE, "libnldpc_b200.so")   # env override: kernel experiments only

NLDPC_OUT_NONE, NLDPC_OUT_ALL, NLDPC_OUT_LAST = 0, 1, 2
NLDPC_LLR_F16, NLDPC_LLR_Q8 = 1, 2      # nldpc_neural_decode_host_narrow input formats
NLDPC_DEC_SP, NLDPC_DEC_MS, NLDPC_DEC_QMS = 0, 1, 2

_lib = None
_lock = threading.Lock()


class NldpcError(RuntimeError):
    pass


class BoostedCfg(ctypes.Structure):
    _fields_ = [("decoder_type", ctypes.c_int32), ("qbit", ctypes.c_int32), ("llr_lo", ctypes.c_float),
                ("llr_hi", ctypes.c_float), ("compute_ucn", ctypes.c_int32), ("ucn_mix", ctypes.c_int32),
                ("llr_init", ctypes.c_void_p), ("xin_init", ctypes.c_void_p), ("xin_out", ctypes.c_void_p),
                ("app_init", ctypes.c_void_p), ("train_dump", ctypes.c_void_p), ("train_dump_bytes", ctypes.c_size_t),
                ("llr_all", ctypes.c_void_p), ("llr_pitch", ctypes.c_int32)]


def build(verbose=False):
    """Compile the CUDA sources in-tree with nvcc for sm_100a (csrc/Makefile)."""
    import subprocess
    cmd = ["make", "-j8", "-C", os.path.join(_HERE, "csrc")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise NldpcError("building libnldpc_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NldpcError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                                 "(nvcc, sm_100a). There is no CPU / PyTorch fallback for this path.")
            L = ctypes.CDLL(LIB_PATH)
            vp, ci = ctypes.c_void_p, ctypes.c_int
            L.nldpc_last_error.restype = ctypes.c_char_p
            L.nldpc_abi_version.restype = ci
            L.nldpc_graph_create.restype = ci
            L.nldpc_graph_create.argtypes = [vp, ci, ci, ci, ci, ctypes.POINTER(vp)]
            L.nldpc_graph_destroy.restype = None
            L.nldpc_graph_destroy.argtypes = [vp]
            L.nldpc_graph_info.restype = ci
            L.nldpc_graph_info.argtypes = [vp, ctypes.POINTER(ctypes.c_int32)]
            L.nldpc_neural_forward.restype = ci
            L.nldpc_neural_forward.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, ci, vp, vp]
            L.nldpc_neural_decode_host.restype = ci
            L.nldpc_neural_decode_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, vp, ci, vp]
            L.nldpc_neural_decode_host_narrow.restype = ci
            L.nldpc_neural_decode_host_narrow.argtypes = [vp, vp, ci, ctypes.c_float, vp, vp, ci, ci, ci, vp, ci, vp]
            L.nldpc_backward_workspace_bytes.restype = ctypes.c_size_t
            L.nldpc_backward_workspace_bytes.argtypes = [vp, ci, ci, ci]
            L.nldpc_neural_backward.restype = ci
            L.nldpc_neural_backward.argtypes = [vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, ctypes.c_size_t, ci, vp]
            L.nldpc_neural_forward_train.restype = ci
            L.nldpc_neural_forward_train.argtypes = [vp, vp, vp, vp, ci, ci, vp, vp, ctypes.c_size_t, vp]
            L.nldpc_boosted_backward.restype = ci
            L.nldpc_boosted_backward.argtypes = [vp, ctypes.POINTER(BoostedCfg), vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp,
                                                 ctypes.c_size_t, ci, vp]
            L.nldpc_boosted_forward.restype = ci
            L.nldpc_boosted_forward.argtypes = [vp, ctypes.POINTER(BoostedCfg), vp, vp, vp, vp, ci, ci, ci, vp, ci, vp, vp, vp]
            L.nldpc_boosted_dump_format.restype = ci
            L.nldpc_boosted_dump_format.argtypes = [vp, ctypes.POINTER(BoostedCfg), ci, ci, ci]
            L.nldpc_boosted_train_workspace_bytes.restype = ctypes.c_size_t
            L.nldpc_boosted_train_workspace_bytes.argtypes = [vp, ctypes.POINTER(BoostedCfg), ci, ci, ci, ci]
            L.nldpc_boosted_train_forward.restype = ci
            L.nldpc_boosted_train_forward.argtypes = [vp, ctypes.POINTER(BoostedCfg), vp, vp, vp, ci, ci, vp, vp, ctypes.c_float, vp, vp,
                                                      ctypes.c_size_t, vp]
            L.nldpc_boosted_train_backward.restype = ci
            L.nldpc_boosted_train_backward.argtypes = [vp, ctypes.POINTER(BoostedCfg), vp, vp, vp, ci, ci, vp, vp, vp, ctypes.c_size_t, vp]
            L.nldpc_pack_labels.restype = ci
            L.nldpc_pack_labels.argtypes = [vp, ctypes.c_size_t, ci, vp, vp]
            L.nldpc_multi_iter_bce.restype = ci
            L.nldpc_multi_iter_bce.argtypes = [vp, vp, vp, ci, ctypes.c_size_t, vp, vp, vp]
            L.nldpc_multi_iter_bce_grad.restype = ci
            L.nldpc_multi_iter_bce_grad.argtypes = [vp, vp, vp, vp, ci, ctypes.c_size_t, vp, vp]
            L.nldpc_boosted_decode_host_q8.restype = ci
            L.nldpc_boosted_decode_host_q8.argtypes = [vp, ctypes.POINTER(BoostedCfg), vp, ctypes.c_float, vp, vp, vp, ci, ci, ci, vp, ci, vp]
            L.nldpc_clip_adam_clamp.restype = ci
            L.nldpc_clip_adam_clamp.argtypes = [vp, vp, vp, vp, vp, ci, ci, ctypes.c_float, ctypes.c_float, ctypes.c_double, vp,
                                                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_float, vp]
            L.nldpc_count_errors.restype = ci
            L.nldpc_count_errors.argtypes = [vp, ctypes.c_size_t, vp, ci, ci, ci, vp, vp]
            L.nldpc_count_errors_packed.restype = ci
            L.nldpc_count_errors_packed.argtypes = [vp, ctypes.c_size_t, vp, ci, ci, ci, vp, vp]
            _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().nldpc_last_error()
        raise NldpcError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


class GraphHandle:
    """Owns one nldpc_graph_t (device-resident Tanner tables) — the replacement of the dense
    ConnectingMatrixTorch tensors on the hot path."""

    def __init__(self, basegraph, Z, device_index):
        bg = np.ascontiguousarray(np.asarray(basegraph), dtype=np.int32)
        self.M, self.N = (int(v) for v in bg.shape)
        self.Z = int(Z)
        self.device_index = int(device_index)
        h = ctypes.c_void_p()
        check(lib().nldpc_graph_create(bg.ctypes.data_as(ctypes.c_void_p), self.M, self.N, self.Z, self.device_index,
                                       ctypes.byref(h)), "nldpc_graph_create")
        self.ptr = h
        info = (ctypes.c_int32 * 8)()
        check(lib().nldpc_graph_info(self.ptr, info), "nldpc_graph_info")
        self.E, self.S = int(info[3]), int(info[4])
        self.cw_per_cta, self.threads_per_cta, self.specialised = int(info[5]), int(info[6]), bool(info[7])
        self.NZ = self.N * self.Z
        self.hard_bytes = (self.NZ + 7) // 8

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and _lib is not None:
                _lib.nldpc_graph_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


_graphs = {}      # (device_index, M, N, Z, bytes) -> GraphHandle
_graph_ids = []   # id -> GraphHandle (ids are what the torch.library ops carry)


def graph_id_for(basegraph, Z, device_index):
    bg = np.ascontiguousarray(np.asarray(basegraph), dtype=np.int32)
    key = (int(device_index), bg.shape, int(Z), bg.tobytes())
    with _lock:
        gid = _graphs.get(key)
    if gid is None:
        h = GraphHandle(bg, Z, device_index)
        with _lock:
            _graph_ids.append(h)
            gid = len(_graph_ids) - 1
            _graphs[key] = gid
    return gid


def graph_by_id(gid):
    return _graph_ids[gid]
