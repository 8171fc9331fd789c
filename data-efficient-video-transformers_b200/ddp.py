"""Clip-batch data parallelism: one process per GPU, replicated parameters, gradients all-reduced over
NCCL (NVLink 5 / NVSwitch) in flat fp32 buckets that are launched from autograd hooks while the rest of
the backward pass is still running (the reference has no multi-GPU path at all: pl.Trainer(gpus=1),
src/main.py:87-88; the oracle for this module is "single process on the concatenated batch").

Parameters' ``.grad`` tensors are views into the flat buckets, so there is no gather/scatter copy: the
wgrad kernels' outputs are accumulated by autograd straight into the communication buffer.  Buckets are
filled in reverse registration order (the order backward produces gradients).  Works with any
torch.distributed backend (NCCL on GPUs; gloo in the CPU tests of the bucketing logic).
"""
import os
import weakref

import torch
import torch.distributed as dist

# id(parameter) -> (weakref to the parameter, weakref to its reducer) for reducers created with direct=True
_DIRECT = {}


def direct_target(p):
    """If ``p``'s gradient is a view into a bucket of a live ``GradBucketReducer(direct=True)``: ``(grad view, done)``,
    where a backward kernel may ACCUMULATE (atomics / +=) straight into ``grad view`` and must call ``done()`` once its
    launches are issued, instead of handing a temporary to autograd (which would cost a zero-fill, a temporary and an
    ``add_`` launch per parameter).  Otherwise ``None``."""
    ent = _DIRECT.get(id(p))
    if ent is None:
        return None
    pr, rr = ent[0](), ent[1]()
    if pr is not p or rr is None or p.grad is None:
        return None
    return p.grad, (lambda: rr._on_direct(p))


def note_use(p):
    """Called by an autograd Function's FORWARD for every parameter whose gradient its backward will write through a
    direct sink: the reducer counts the uses, and a parameter only counts as complete once every use has reported
    ``done()``.  A module applied several times per step (``SimpleTransformer.ptn_shared`` runs ``transformer_encoder0``
    once per expert, src/models/transformer.py:84-104) would otherwise arm its bucket's all-reduce on the FIRST
    contribution while later ones are still accumulating into the same flat buffer."""
    ent = _DIRECT.get(id(p))
    if ent is None:
        return
    pr, rr = ent[0](), ent[1]()
    if pr is p and rr is not None:
        rr._uses[id(p)] = rr._uses.get(id(p), 0) + 1


class GradBucketReducer:
    def __init__(self, params, bucket_bytes=32 << 20, process_group=None, average=True, direct=True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # measurement knob (bench.py records it): 1 = never issue the all-reduces, i.e. N independent replicas in lockstep - the
        # difference to a normal run is the whole cost of gradient communication, interference with backward included
        self.dry_run = os.environ.get("TVT_DDP_DRY_RUN", "0") == "1"
        self.buckets = []          # list of dict(flat, params, pending, handle)
        self._index = {}
        self._uses = {}            # id(param) -> direct-sink uses registered by forward and not yet reported done
        cur, cur_bytes = [], 0
        for p in reversed(self.params):                      # backward order
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._make_bucket(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._make_bucket(cur)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.direct = direct
        if direct:
            me = weakref.ref(self)
            for p in self.params:
                _DIRECT[id(p)] = (weakref.ref(p), me)

    ALIGN = 8   # elements: every view starts 32-byte aligned (16 bytes for the bf16 planes laid out alike)

    def _make_bucket(self, ps):
        offsets, n = [], 0
        for p in ps:
            offsets.append(n)
            n += -(-p.numel() // self.ALIGN) * self.ALIGN
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        b = {"flat": flat, "params": list(ps), "offsets": offsets, "pending": len(ps), "handle": None, "seen": set()}
        for p, off in zip(ps, offsets):
            if p.dtype != torch.float32:
                raise TypeError("GradBucketReducer expects fp32 master parameters")
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._index[id(p)] = b
        self.buckets.append(b)

    def zero_grad(self):
        """Zero every bucket (one memset per bucket) and re-arm the hooks; call before each backward."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["handle"] = None
            b["seen"].clear()
        self._uses.clear()

    def _on_direct(self, p):
        """One direct-sink use of ``p`` has issued its kernels; the parameter completes when all registered uses have."""
        left = self._uses.get(id(p), 1) - 1
        self._uses[id(p)] = max(left, 0)
        if left <= 0:
            self._on_grad(p)

    def _on_grad(self, p):
        b = self._index[id(p)]
        if id(p) in b["seen"]:          # a further backward without zero_grad(): the bucket is already accounted for
            return
        b["seen"].add(id(p))
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1 and not self.dry_run:
            # async: NCCL's stream waits on the producer stream, backward keeps going on the compute stream
            b["handle"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket, reduce the ones whose hooks never fired (unused params), average
        (unless ``average=False``: FlatOptimizer fuses the 1 / world scaling into its step kernel)."""
        for b in self.buckets:
            if self.world > 1 and not self.dry_run:
                if b["handle"] is None:
                    b["handle"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                b["handle"].wait()
                if self.average:
                    b["flat"].mul_(1.0 / self.world)

    def remove(self):
        for h in self._hooks:
            h.remove()
        for p in self.params:
            ent = _DIRECT.get(id(p))
            if ent is not None and ent[1]() is self:
                del _DIRECT[id(p)]


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns (rank, local, world)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world
