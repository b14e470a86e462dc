"""`torch_scatter.scatter` subset used at Ob_propagation.py:227 (test infrastructure)."""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    if dim < 0:
        dim += src.dim()
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    shape = list(src.shape)
    shape[dim] = n
    view = [1] * src.dim()
    view[dim] = -1
    idx = index.view(view).expand_as(src)
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    if reduce in ("sum", "add"):
        return res.scatter_add(dim, idx, src)
    if reduce == "mean":
        return res.scatter_reduce(dim, idx, src, reduce="mean", include_self=False)
    if reduce == "max":
        return res.scatter_reduce(dim, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


def gather_csr(*a, **k):  # imported, never called (Ob_propagation.py:14)
    raise NotImplementedError


def segment_csr(*a, **k):  # imported, never called (Ob_propagation.py:14)
    raise NotImplementedError
