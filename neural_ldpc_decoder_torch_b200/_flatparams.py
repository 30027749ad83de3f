"""Flat parameter storage for the decoders' (tiny) weight vectors.

The decode launches want the weights as `[T, ·]` rows.  Stacking / folding them per call costs more host time than a
batch-1024 decode, and caching the stacked COPY would go stale whenever a caller updates a parameter in a way that does not
bump `Tensor._version` (`p.data.clamp_()` as the reference's `_apply_constraints` does, a kernel writing through a raw pointer).
So the modules keep their parameters as slices of ONE flat fp32 vector and read the rows LIVE: as a strided view when the
rows are consecutive, else through a cached gather index (one launch).  What is cached is only the layout (keyed by the
parameters' data pointers), never a value.  Anything that breaks the layout (`p.data = other`) falls back to per-call
stacking, which is always correct."""
import torch


def flatten_(params):
    """Re-point `.data` of every parameter at consecutive slices of one new flat fp32 vector on the parameters' device
    (values preserved; names, shapes and Parameter objects untouched).  Returns the flat vector (None if not applicable)."""
    params = list(params)
    if not params:
        return None
    dev = params[0].device
    if any(p.device != dev or p.dtype != torch.float32 for p in params):
        return None
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        off = 0
        for p in params:
            n = p.numel()
            p.data = flat[off:off + n].view(p.shape)
            off += n
    return flat


def storage_base(params, device):
    """1-D fp32 tensor over the WHOLE storage shared by all `params` (they must be contiguous fp32 tensors of one storage on
    `device`), else None."""
    params = list(params)
    if not params:
        return None
    st = params[0].untyped_storage()
    sp = st.data_ptr()
    for p in params:
        if p.device != device or p.dtype != torch.float32 or not p.is_contiguous() or p.untyped_storage().data_ptr() != sp:
            return None
    return torch.empty(0, dtype=torch.float32, device=device).set_(st)


def element_offset(p, base):
    return (p.data_ptr() - base.data_ptr()) // 4
