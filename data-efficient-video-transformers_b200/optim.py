"""Flat-bucket optimizer: AdamW / SGD-momentum / Adagrad applied by ONE kernel launch per
gradient bucket.  Parameters are re-pointed to views of a flat fp32 buffer laid out exactly like the
``GradBucketReducer``'s gradient buckets, so the step is a perfectly coalesced pass that also
  * applies the data-parallel gradient averaging (1 / world), and
  * writes the bf16 operand planes the next step's GEMMs read (registered in the modules' ``ops.Mode`` caches),
replacing torch's multi-tensor optimizer, the bucket scaling pass and ~60 per-weight conversion launches.
Update rules and defaults are torch.optim.AdamW's / SGD's / Adagrad's (the reference's configure_optimizers:
src/models/frame_transformer.py:123-134, src/models/transformer.py:58-64)."""
import weakref

import torch

from . import capi, ops


class FlatOptimizer:
    def __init__(self, reducer, modes=(), kind="adamw", lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, momentum=0.0):
        if kind not in ("adamw", "sgd", "adagrad"):
            raise ValueError("kind must be 'adamw', 'sgd' or 'adagrad'")
        self.reducer, self.modes = reducer, list(modes)
        self.kind = {"adamw": 0, "sgd": 1, "adagrad": 2}[kind]
        self.lr, self.betas, self.eps, self.weight_decay, self.momentum = lr, betas, eps, weight_decay, momentum
        self.step_count = 0
        self.step_dev = None        # device int64 step count (set by hostapi.GraphedTrainStep): the kernel reads it instead of step_count
        self.buckets = []
        want_lo = any(m.fp32 for m in self.modes)
        for b in reducer.buckets:
            flat_g = b["flat"]
            flat_p = torch.zeros_like(flat_g)
            views = []
            for p, off in zip(b["params"], b["offsets"]):        # same (aligned) layout as the gradient bucket
                n = p.numel()
                flat_p[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat_p[off:off + n].view_as(p)           # the parameter now lives in the flat buffer
                views.append((p, off, n))
            hi = torch.empty(flat_g.numel(), dtype=torch.bfloat16, device=flat_g.device) if self.modes else None
            lo = torch.empty_like(hi) if (hi is not None and want_lo) else None
            self.buckets.append({"p": flat_p, "g": flat_g, "m": torch.zeros_like(flat_g),
                                 "v": torch.zeros_like(flat_g) if self.kind != 1 else None, "hi": hi, "lo": lo, "views": views})

    def zero_grad(self):
        self.reducer.zero_grad()

    @torch.no_grad()
    def step(self):
        """Call after ``reducer.finish(average=False)``-style reduction: the 1 / world averaging is fused here."""
        self.step_count += 1
        stream = torch.cuda.current_stream().cuda_stream
        scale = 1.0 / self.reducer.world if (self.reducer.world > 1 and not self.reducer.average) else 1.0
        for b in self.buckets:
            a = capi.OptimStepArgs()
            a.p, a.g, a.m = b["p"].data_ptr(), b["g"].data_ptr(), b["m"].data_ptr()
            a.v = b["v"].data_ptr() if b["v"] is not None else None
            a.p_hi = b["hi"].data_ptr() if b["hi"] is not None else None
            a.p_lo = b["lo"].data_ptr() if b["lo"] is not None else None
            a.n, a.kind, a.step = b["p"].numel(), self.kind, self.step_count
            a.step_dev = self.step_dev.data_ptr() if self.step_dev is not None else None
            a.lr, a.beta1, a.beta2, a.eps = self.lr, self.betas[0], self.betas[1], self.eps
            a.weight_decay, a.momentum, a.grad_scale = self.weight_decay, self.momentum, scale
            capi.call("tvt_optim_step", a, stream)
            if b["hi"] is not None:
                for p, off, n in b["views"]:
                    if p.dim() != 2:
                        continue                                   # only matrices are GEMM operands
                    hi = b["hi"][off:off + n].view_as(p)
                    lo = b["lo"][off:off + n].view_as(p) if b["lo"] is not None else None
                    for m in self.modes:
                        m._wcache[id(p)] = (weakref.ref(p), p._version, hi, lo if m.fp32 else None)
