"""Optimiser for the drop-in module path: torch.optim.Adam semantics (code/Raindrop.py:256: lr 1e-4,
betas (0.9, 0.999), eps 1e-8, no weight decay) as ONE kernel launch over a flat parameter bucket.

`torch.optim.Adam(model.parameters())` walks 64 tensors on the host and issues a foreach kernel group per
moment, and autograd visits 34 leaves per backward; at the reference's batch sizes that costs more host time than
the whole forward.  `FlatAdam(model)` re-homes the 34 tensors that actually receive gradient (SURVEY.md section
8a18) as views of one flat fp32 leaf `flat_p`, lets the backward (rd_raindrop_v2_bwd) write their gradients into a
second flat buffer that IS `flat_p.grad` (and of which every `param.grad` is a view), and steps with `rd_adam_step`.
The other 30 parameters never get a gradient in the reference either, so Adam would skip them -- same trajectory.

    opt = FlatAdam(model, lr=1e-4)          # instead of torch.optim.Adam(model.parameters(), lr=1e-4)
    loss.backward(); opt.step(); opt.zero_grad()

It is a `torch.optim.Optimizer`: `param_groups[0]["lr"]` is honoured every step, so
`ReduceLROnPlateau` (code/Raindrop.py:257-259) works unchanged.  Gradients are OVERWRITTEN by every backward
(`zero_grad()` is a no-op): the reference never accumulates gradients over several backwards; for that use
torch.optim.Adam, which the module supports as well.
"""
import weakref

import torch

from . import lib as L


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = model.used_parameters()
        dev = params[0].device
        if dev.type != "cuda":
            raise L.RaindropB200Error("FlatAdam needs the model on a CUDA device (call model.cuda() first)")
        self.lib = L.load()
        self.model = model
        self.offsets, total = [], 0
        for p in params:               # same order as the backward's gradient bucket, 16-byte aligned pieces
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat_p = torch.nn.Parameter(torch.zeros(total, dtype=torch.float32, device=dev))
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(params, self.offsets):
            view = self.flat_p.data[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view(p.shape)      # every .grad is a window of the bucket
        self.flat_p.grad = self.flat_g
        super().__init__([self.flat_p], dict(lr=lr, betas=betas, eps=eps))
        self.exp_avg = torch.zeros_like(self.flat_g)
        self.exp_avg_sq = torch.zeros_like(self.flat_g)
        self.step_count = torch.zeros(2, dtype=torch.int64, device=dev)      # {count, ticket}
        self._params = params
        self.grads_ready = False
        model._flat_optim = weakref.ref(self)
        model.__dict__.pop("_used_params", None)

    def zero_grad(self, set_to_none=True):
        """Nothing to do: the backward overwrites the whole bucket."""
        self.grads_ready = False

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        if not self.grads_ready:
            # the gradients did not come through the flat fast path (e.g. the general autograd path ran while a
            # forward was pending): gather whatever .grad tensors the parameters hold
            for p, off in zip(self._params, self.offsets):
                n = p.numel()
                dst = self.flat_g[off:off + n]
                if p.grad is None:
                    dst.zero_()
                elif p.grad.data_ptr() != dst.data_ptr():
                    dst.copy_(p.grad.reshape(-1))
        g = self.param_groups[0]
        L.check(self.lib.rd_adam_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.flat_p.numel(), float(g["lr"]), None,
                                      float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(grad_scale),
                                      self.step_count.data_ptr(), L.stream_ptr(self.flat_p.device)), "rd_adam_step")
        return loss
