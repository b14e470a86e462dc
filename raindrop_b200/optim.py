"""Optimiser for the drop-in module path: torch.optim.Adam semantics (code/Raindrop.py:256: lr 1e-4,
betas (0.9, 0.999), eps 1e-8, no weight decay) as ONE kernel launch over a flat parameter bucket.

`torch.optim.Adam(model.parameters())` walks 64 tensors on the host and issues a foreach kernel group per
moment; at the reference's batch sizes that costs more host time than the whole forward.  `FlatAdam(model)`
re-homes the 34 tensors that actually receive gradient as views of one flat fp32
buffer, lets the backward (rd_raindrop_v2_bwd) write their gradients into a second flat buffer, and steps
with `rd_adam_step`.  The other 30 parameters never get a gradient in the reference either, so Adam would
skip them (grad is None) -- same trajectory.

    opt = FlatAdam(model, lr=1e-4)          # instead of torch.optim.Adam(model.parameters(), lr=1e-4)
    loss.backward(); opt.step(); opt.zero_grad()

It is a `torch.optim.Optimizer`: `param_groups[0]["lr"]` is honoured every step, so
`ReduceLROnPlateau` (code/Raindrop.py:257-259) works unchanged.
"""
import torch

from . import lib as L


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        params = model.used_parameters()
        dev = params[0].device
        if dev.type != "cuda":
            raise L.RaindropB200Error("FlatAdam needs the model on a CUDA device (call model.cuda() first)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.lib = L.load()
        self.model = model
        total = sum(p.numel() for p in params)
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:               # tightly packed, same order as the backward's gradient bucket
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_count = torch.zeros(2, dtype=torch.int64, device=dev)      # {count, ticket}
        self._params = params
        model._flat_grad_static = self.flat_g       # the backward writes straight into this bucket
        model.__dict__.pop("_used_params", None)

    def _grads_in_bucket(self):
        lo, hi = self.flat_g.data_ptr(), self.flat_g.data_ptr() + 4 * self.flat_g.numel()
        off = 0
        for p in self._params:
            g = p.grad
            if g is None or g.data_ptr() != lo + 4 * off or not g.is_contiguous():
                return False
            off += p.numel()
        return lo + 4 * off == hi

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        if all(p.grad is None for p in self._params):
            return loss
        if not self._grads_in_bucket():      # e.g. gradient accumulation made autograd copy: gather once
            off = 0
            for p in self._params:
                n = p.numel()
                if p.grad is None:
                    self.flat_g[off:off + n].zero_()
                else:
                    self.flat_g[off:off + n].copy_(p.grad.reshape(-1))
                off += n
        g = self.param_groups[0]
        L.check(self.lib.rd_adam_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), self.flat_p.numel(), float(g["lr"]), None,
                                      float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(grad_scale),
                                      self.step_count.data_ptr(), L.stream_ptr(self.flat_p.device)), "rd_adam_step")
        return loss
