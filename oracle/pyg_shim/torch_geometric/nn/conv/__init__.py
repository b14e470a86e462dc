"""`torch_geometric.nn.conv.MessagePassing` subset (test infrastructure).

Only what Ob_propagation.py:114 and transformer_conv.py:158 reach: dense `edge_index`
[2, E], flow source_to_target (x_j <- edge_index[0], x_i <- edge_index[1],
index = edge_index[1]), node_dim = 0, aggr = 'add', identity update.
"""
import inspect

import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        self._msg_params = None

    def propagate(self, edge_index, size=None, **kwargs):
        if self._msg_params is None:
            self._msg_params = list(inspect.signature(self.message).parameters)
        assert torch.is_tensor(edge_index) and edge_index.dim() == 2 and edge_index.size(0) == 2
        src_idx, tgt_idx = edge_index[0], edge_index[1]
        n_tgt = None
        call = {}
        for name in self._msg_params:
            if name in ("size_i", "size_j"):
                continue
            if name.endswith("_i") or name.endswith("_j"):
                base = name[:-2]
                data = kwargs[base]
                pick = 1 if name.endswith("_i") else 0
                if isinstance(data, (tuple, list)):
                    if n_tgt is None and data[1] is not None:
                        n_tgt = data[1].size(self.node_dim)
                    data = data[pick]
                elif torch.is_tensor(data) and n_tgt is None:
                    n_tgt = data.size(self.node_dim)
                if torch.is_tensor(data):
                    data = data.index_select(self.node_dim, tgt_idx if pick == 1 else src_idx)
                call[name] = data
        if n_tgt is None:
            n_tgt = int(tgt_idx.max()) + 1
        for name in self._msg_params:
            if name in call:
                continue
            if name == "index":
                call[name] = tgt_idx
            elif name == "ptr":
                call[name] = None
            elif name == "size_i":
                call[name] = n_tgt
            elif name == "size_j":
                call[name] = n_tgt
            else:
                call[name] = kwargs.get(name, None)
        out = self.message(**call)
        out = self.aggregate(out, index=tgt_idx, ptr=None, dim_size=n_tgt)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs
