"""Evaluation path: the step after the hot path.

The reference's ``validation_step`` (src/models/transformer.py:146-158) appends ``sigmoid(logits)``, the raw logits and
``target.int()`` to Python lists, and its callback (src/callbacks/callbacks.py:34-45) concatenates them, moves them to the
host and thresholds them once per threshold in ``[0, 0.1 ... 0.8]`` for sklearn.  Here

* ``EvalBuffer`` is ONE preallocated device buffer per quantity that every batch is written into by a single kernel
  launch (``tvt_eval_readout``: sigmoid, label cast, all thresholds as a bit mask, top-1 index); the sklearn metrics
  stay on the CPU and read one contiguous ``[N, C]`` array per quantity;
* ``GraphedForward`` captures the inference-only forward of a module (no saved activations, dropout off) in a CUDA
  graph, so the launch-bound small configurations (BASELINE C1-C3) replay ~100 kernels with one launch.
"""
import torch

from .. import ops

REFERENCE_THRESHOLDS = (0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8)   # callbacks.py:37


class EvalBuffer:
    """Running evaluation buffers on the device: ``probs`` [N, C] fp32, ``labels`` [N, C] int32, ``pred_bits`` [N, C]
    (bit k: prob > thresholds[k]) and ``top1`` [N]; ``append`` is one kernel launch per batch."""

    def __init__(self, capacity, n_classes, thresholds=REFERENCE_THRESHOLDS, device="cuda"):
        if len(thresholds) > 16:
            raise ValueError("EvalBuffer: at most 16 thresholds")
        self.capacity, self.n_classes, self.thresholds = int(capacity), int(n_classes), tuple(float(t) for t in thresholds)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError("EvalBuffer lives on a CUDA device: this path has no CPU implementation")
        self._probs = torch.zeros(self.capacity, self.n_classes, dtype=torch.float32, device=dev)
        self._labels = torch.zeros(self.capacity, self.n_classes, dtype=torch.int32, device=dev)
        self._bits = torch.zeros(self.capacity, self.n_classes, dtype=torch.int16, device=dev)
        self._top1 = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.rows = 0

    def reset(self):
        self.rows = 0

    def append(self, logits, target=None):
        """logits [B, C] (any float dtype), target [B, C] or [B, 1, C] float (the loader's layout) or None."""
        logits = logits.detach().reshape(-1, self.n_classes).float().contiguous()
        if target is not None:
            target = target.detach().reshape(-1, self.n_classes)
            if target.dtype not in (torch.float32, torch.float64):
                target = target.float()
            if target.shape[0] != logits.shape[0]:
                raise ValueError(f"EvalBuffer.append: {logits.shape[0]} logit rows but {target.shape[0]} target rows")
            target = target.to(logits.device)
        if self.rows + logits.shape[0] > self.capacity:
            raise ValueError(f"EvalBuffer.append: {self.rows} + {logits.shape[0]} rows exceed the capacity {self.capacity}")
        ops.eval_readout(logits, target, self._probs, self._labels, self._bits, self._top1, self.rows, self.thresholds)
        self.rows += logits.shape[0]

    # views of the filled part
    @property
    def probs(self):
        return self._probs[: self.rows]

    @property
    def labels(self):
        return self._labels[: self.rows]

    @property
    def top1(self):
        return self._top1[: self.rows]

    def predictions(self, threshold):
        """(probs > threshold) as int32 [N, C] for one of the configured thresholds."""
        k = self.thresholds.index(float(threshold))
        return ((self._bits[: self.rows].to(torch.int32) >> k) & 1)

    def to_host(self):
        """The arrays the reference's callback hands to sklearn, each with ONE device-to-host copy."""
        return {"probs": self.probs.cpu().numpy(), "labels": self.labels.cpu().numpy(), "top1": self.top1.cpu().numpy(),
                "pred_bits": self._bits[: self.rows].cpu().numpy()}


class GraphedForward:
    """CUDA-graph replay of ``fn(*inputs)`` for fixed input shapes, inference only (``torch.no_grad``; put the module in
    ``eval()`` first - a captured graph would replay the same dropout seeds).  ``fn`` may return a tensor or a tuple /
    list of tensors; the returned tensors are static buffers overwritten by the next call."""

    def __init__(self, fn, example_inputs, warmup=2, module=None):
        """``module`` (default: ``fn.__self__`` when ``fn`` is a bound method of an nn.Module): the module whose weights the
        graph reads.  The captured kernels hold the RAW ADDRESSES of the bf16 weight planes cached in ``ops.Mode`` at capture
        time; a weight update that re-splits a parameter into new planes (a plain torch optimizer step, ``load_state_dict``:
        anything that bumps ``Parameter._version`` outside ``optim.FlatOptimizer``, which refreshes its planes in place) would
        leave the graph replaying stale or freed memory.  ``__call__`` therefore compares every parameter's version and
        storage address with the ones recorded at capture and RE-CAPTURES the graph when any of them has changed."""
        self.fn = fn
        self.module = module if module is not None else getattr(fn, "__self__", None)
        if not isinstance(self.module, torch.nn.Module):
            self.module = None
        self.warmup = warmup
        self.captures = 0
        self.static_in = [x.clone() for x in example_inputs]
        self._capture()

    def _weights_key(self):
        if self.module is None:
            return None
        return tuple((id(p), p._version, p.data_ptr()) for p in self.module.parameters())

    def _capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):     # lazy initialisation (kernel attributes, weight plane caches) stays out of the graph
                self.fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.fn(*self.static_in)
        self._planes = self._plane_key()
        self._key = self._weights_key()
        self.captures += 1

    def _plane_key(self):
        """Addresses of the bf16 operand planes the captured kernels read (every ``ops.Mode`` found on the module tree)."""
        if self.module is None:
            return None
        modes = {id(m.mode): m.mode for m in self.module.modules() if isinstance(getattr(m, "mode", None), ops.Mode)}
        key = []
        for mode in modes.values():
            for pid, ent in mode._wcache.items():
                key.append((str(pid), ent[2].data_ptr(), ent[3].data_ptr() if ent[3] is not None else 0))
        return tuple(sorted(key))

    def _stale(self):
        if self.module is None or self._key == self._weights_key():
            return False
        # the weights changed.  FlatOptimizer rewrites the cached planes in place and re-registers them under the new version:
        # the addresses the graph holds are still the live ones, so only a change of a plane ADDRESS (or an entry that no
        # longer matches its parameter's version, i.e. one the next eager call would re-split) invalidates the capture.
        now = self._weights_key()
        if len(now) != len(self._key) or any(a[0] != b[0] or a[2] != b[2] for a, b in zip(now, self._key)):
            return True                      # a parameter moved (biases, LayerNorm and head weights are read in place)
        modes = {id(m.mode): m.mode for m in self.module.modules() if isinstance(getattr(m, "mode", None), ops.Mode)}
        if any(mode.any_stale() for mode in modes.values()):
            return True
        if self._planes != self._plane_key():
            return True
        self._key = self._weights_key()
        return False

    def __call__(self, *inputs):
        if self._stale():
            self._capture()
        if len(inputs) != len(self.static_in):
            raise ValueError(f"GraphedForward: expected {len(self.static_in)} inputs, got {len(inputs)}")
        for dst, src in zip(self.static_in, inputs):
            if dst.shape != src.shape:
                raise ValueError(f"GraphedForward: captured for input shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """Whole-step CUDA graph of a TRAINING step: ``step_fn(*inputs) -> loss`` (zero_grad, forward, backward, bucket
    all-reduces, optimizer step — everything the step launches) is captured once per distinct set of input buffers and
    replayed with a single launch.  The launch-bound BASELINE configs (C1: ~100 kernels of a few microseconds each) spend
    most of a step in launch gaps; a replay has none.

    What changes from step to step cannot be a kernel argument inside a graph, so it lives in device memory:
      * dropout masks: ``tvt_set_seed_source`` registers a device step counter; every mask-drawing kernel folds its hash
        into the (baked) seed, forward and backward alike; the first node of the graph advances the counter;
      * the optimizer's bias-correction step count: ``FlatOptimizer.step_dev`` (the element next to the seed counter).
    The gradient buckets, optimizer state and bf16 weight planes are persistent buffers updated in place, so their
    addresses are stable by construction (``optim.FlatOptimizer`` is required: a torch optimizer would re-split weights).
    Replaying with inputs at new addresses captures another graph that shares the first one's memory pool; the returned loss
    is a static tensor overwritten by the next replay.  Eager calls of ``step_fn`` stay valid between replays (they fold the
    same device counter)."""

    def __init__(self, step_fn, optimizer, example_inputs=None, warmup=3):
        from .. import capi
        self.step_fn, self.opt = step_fn, optimizer
        dev = optimizer.buckets[0]["p"].device
        self.counters = torch.zeros(2, dtype=torch.int64, device=dev)          # [dropout step counter, optimizer step count]
        self.counters[1] = optimizer.step_count
        optimizer.step_dev = self.counters[1:]
        capi.set_seed_source(self.counters.data_ptr())
        self.graphs = {}
        self.pool = None
        self.warmup = warmup
        self.kernels_per_replay = 0
        self._warm = False
        if example_inputs is not None:
            self._graph_for(example_inputs)

    @staticmethod
    def _flat(inputs):
        out = []
        for x in inputs:
            out.extend(x) if isinstance(x, (list, tuple)) else out.append(x)
        return out

    def _advance(self):
        from .. import capi
        capi.step_counter_advance(self.counters.data_ptr(), 2, torch.cuda.current_stream().cuda_stream)

    def _graph_for(self, inputs):
        from .. import capi
        key = tuple(t.data_ptr() for t in self._flat(inputs))
        ent = self.graphs.get(key)
        if ent is not None:
            return ent
        if not self._warm:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup):       # lazy initialisation (kernel attributes, weight planes, autograd, NCCL) outside the graph
                    self._advance()
                    self.step_fn(*inputs)
            torch.cuda.current_stream().wait_stream(side)
            self._warm = True
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        l0 = capi.launches
        from .. import functions
        self.capture_seed_counter = functions._seed_counter[0]     # the baked host seeds derive from this (tests replay it eagerly)
        with torch.cuda.graph(g, pool=self.pool):
            self._advance()
            loss = self.step_fn(*inputs)
        self.kernels_per_replay = capi.launches - l0
        if self.pool is None:
            self.pool = g.pool()
        ent = (g, loss, inputs, self.capture_seed_counter)   # the inputs are kept alive: the graph reads their addresses
        self.graphs[key] = ent
        return ent

    def prepare(self, *inputs):
        """Capture the graph for this set of input buffers now (e.g. before a timed region).  Returns the value the host
        seed counter (functions._seed_counter) had when the graph's seeds were drawn."""
        return self._graph_for(inputs)[3]

    def __call__(self, *inputs):
        g, loss = self._graph_for(inputs)[:2]
        g.replay()
        self.opt.step_count += 1
        return loss

    def close(self):
        from .. import capi
        capi.set_seed_source(None)
        self.opt.step_dev = None
        self.graphs.clear()
