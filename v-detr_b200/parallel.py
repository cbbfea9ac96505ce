"""Data-parallel plumbing for the decoder hot path.

The path shards naturally: scenes are independent, so a global batch is split by scene over ranks (one process per
GPU) and the only data-path collective is the bucketed gradient all-reduce that DistributedDataParallel issues
during backward (the reference does the same: main.py:515-517).  BatchNorm statistics stay per GPU (the reference
converts to SyncBatchNorm, main.py:512-514; with 8 scenes x 1024 queries per GPU the local statistics already cover
8192 samples per channel -- stated deviation).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous, balanced range of scene indices owned by `rank` (first `global_batch % world` ranks get one more)."""
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def wrap_data_parallel(module: torch.nn.Module, device=None, find_unused_parameters: bool = False):
    """DistributedDataParallel with the reference's flags (main.py:515-517); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return module
    ids = None if device is None or torch.device(device).type != "cuda" else [torch.device(device).index]
    return torch.nn.parallel.DistributedDataParallel(module, device_ids=ids, find_unused_parameters=find_unused_parameters)


def max_over_ranks(value: float, device=None) -> float:
    """Every multi-GPU time is reported as the maximum over ranks (never wall clock)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class FlatGradAllReduce:
    """ONE all-reduce of all gradients per step (the "single NCCL all-reduce on gradients" of the north star).

    All ``.grad`` tensors are views into one flat buffer, so synchronisation is a single collective on a static
    address -- which also makes the whole step (forward, backward, this all-reduce, optimizer) capturable in one CUDA
    graph.  The decoder's gradients are 46.5 MB fp32; on NVLink 5 that is a few hundred microseconds against a
    ~100 ms step, so overlapping it with backward (what DistributedDataParallel's buckets do) buys nothing here.
    """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._attached = True          # .grad tensors are the views of the flat buffer

    def zero_(self):
        """Before backward.  The gradients are detached from the flat buffer: with ``.grad is None`` autograd hands
        every parameter its freshly computed gradient (no kernel) instead of issuing one ``add_`` per parameter into
        a zeroed view -- ~1000 tiny launches per step for this decoder."""
        for p in self.params:
            p.grad = None
        self._attached = False

    def gather_(self):
        """After backward: copy the gradients into the flat buffer (a few multi-tensor launches) and make ``.grad``
        the views again, so that the all-reduce and the optimizer see static addresses."""
        if self._attached:
            return self.flat
        have_v, have_g = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                have_v.append(v)
                have_g.append(p.grad)
        if len(have_v) < len(self.params):
            self.flat.zero_()
        if have_v:
            torch._foreach_copy_(have_v, have_g)
        for p, v in zip(self.params, self.views):
            p.grad = v
        self._attached = True
        return self.flat

    def sync_(self, average: bool = True):
        """SUM all-reduce of the flat buffer; average=False leaves the 1 / world factor to the optimizer (FlatAdamW folds
        it into its gradient scale: no extra pass over the buffer)."""
        self.gather_()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


class FlatAdamW:
    """torch.optim.AdamW semantics (the reference: optimizer.py:4-26, stepped after clip_grad_norm_ in engine.py:105-108) on
    FLAT buffers: parameters are re-homed as views of one fp32 buffer, gradients are the views of FlatGradAllReduce, the two
    moments are flat, and a step is ONE kernel launch (csrc/optim.cu) instead of ~22 multi-tensor launches over ~1000 small
    tensors.  lr, the step count and the gradient scale live on the device, so the step can sit inside a captured CUDA graph
    and still follow a learning-rate schedule (set_lr) or a clipped gradient norm.

    no_decay(name, param) -> True puts a parameter into the group without weight decay (--filter_biases_wd)."""

    def __init__(self, named_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, no_decay=None):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("FlatAdamW: no trainable parameters")
        dec = [(n, p) for n, p in named if not (no_decay and no_decay(n, p))]
        nod = [(n, p) for n, p in named if no_decay and no_decay(n, p)]
        self.names = [n for n, _ in dec + nod]
        self.params = [p for _, p in dec + nod]
        ref = self.params[0]
        if not ref.is_cuda or any(p.dtype != torch.float32 or p.device != ref.device for p in self.params):
            raise RuntimeError("FlatAdamW needs fp32 CUDA parameters on one device (there is no CPU path)")
        self.n = sum(p.numel() for p in self.params)
        self.n_decay = sum(p.numel() for _, p in dec)
        self.flat_p = torch.empty(self.n, dtype=torch.float32, device=ref.device)
        o = 0
        with torch.no_grad():
            for p in self.params:                      # parameters become views of the flat buffer (same values)
                v = self.flat_p[o:o + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
                o += p.numel()
        self.grads = FlatGradAllReduce(self.params)    # same order: flat_g[i] is the gradient of flat_p[i]
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.lr = torch.tensor([lr], dtype=torch.float32, device=ref.device)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=ref.device)
        self.grad_scale = torch.ones(1, dtype=torch.float32, device=ref.device)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.world_scale = 1.0                         # 1 / world size when the all-reduce is a plain SUM

    def set_lr(self, lr: float):
        self.lr.fill_(lr)

    def clip_grad_norm_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_ (engine.py:105-106) as a device-side scale of the flat gradient; returns the norm."""
        norm = torch.linalg.vector_norm(self.grads.gather_()) * self.world_scale
        torch.clamp(max_norm / (norm + 1e-6), max=1.0, out=self.grad_scale[0])
        return norm

    @torch.no_grad()
    def step(self):
        from . import _C
        self.grads.gather_()
        self.step_t += 1.0
        with torch.cuda.device(self.flat_p.device):
            _C.check(_C.lib().vdetr_adamw_flat(_C.ptr(self.flat_p), _C.ptr(self.grads.flat), _C.ptr(self.exp_avg),
                                               _C.ptr(self.exp_avg_sq), self.n, self.n_decay, _C.ptr(self.lr), _C.ptr(self.step_t),
                                               _C.ptr(self.grad_scale), float(self.world_scale), float(self.betas[0]),
                                               float(self.betas[1]), float(self.eps), float(self.weight_decay), _C.stream_ptr()))

    def zero_grad(self, set_to_none: bool = True):
        self.grads.zero_()

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_t.clone(),
                "lr": self.lr.clone(), "names": list(self.names)}

    def load_state_dict(self, sd):
        if list(sd["names"]) != self.names:
            raise ValueError("FlatAdamW.load_state_dict: parameter order differs")
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_t.copy_(sd["step"]); self.lr.copy_(sd["lr"])
