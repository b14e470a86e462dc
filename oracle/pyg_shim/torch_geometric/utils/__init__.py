"""`torch_geometric.utils.softmax` as used at Ob_propagation.py:195 and transformer_conv.py:201
(test infrastructure).  Segment softmax over dim 0 grouped by `index`:
    out = exp(src - segment_max) / (segment_sum(exp(src - segment_max)) + 1e-16)
"""
import torch


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    assert ptr is None and dim == 0
    n = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    shape = (n,) + tuple(src.shape[1:])
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    seg_max = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
    seg_max = seg_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - seg_max.gather(0, idx)).exp()
    seg_sum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(0, idx, out)
    return out / (seg_sum.gather(0, idx) + 1e-16)
