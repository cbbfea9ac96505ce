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


def flat_shard_range(n: int, rank: int, world: int):
    """[lo, hi) of a flat vector of n floats owned by `rank` in the peer-memory optimizer (csrc/peer.cu): shards tile [0, n),
    every boundary except n itself is a multiple of 4 (the kernels move float4), the last rank takes the tail."""
    n4 = n // 4
    lo = (n4 * rank // world) * 4
    hi = n if rank == world - 1 else (n4 * (rank + 1) // world) * 4
    return lo, hi


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

    def __init__(self, params, flat=None):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device) if flat is None else flat   # (peer memory: PeerGroup.alloc)
        assert self.flat.numel() == n and self.flat.dtype == ref.dtype
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


class _RawCuda:
    """A cudaMalloc allocation owned by libvdetr_b200 (peer memory) exposed to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBuffer:
    """One allocation per rank, every rank's mapped into this process: `local` is this rank's memory as a uint8 tensor,
    `ptrs` a ctypes array of `world` device pointers (entry [rank] = local) in the form the C ABI takes."""

    def __init__(self, local, ptrs, raw):
        self.local, self.ptrs, self._raw = local, ptrs, raw

    def view(self, dtype, numel=None):
        t = self.local.view(dtype)
        return t if numel is None else t[:numel]


class PeerGroup:
    """Peer memory over NVLink for the ranks of one node (csrc/peer.cu): cudaMalloc allocations exchanged as CUDA IPC handles
    through torch.distributed, a flag array for cross-GPU barriers and this rank's epoch counters.  world == 1 works without
    torch.distributed (the exchanges degenerate to local copies), which is how the single-GPU tests cover the kernels."""

    def __init__(self, device):
        from . import _C
        self.device = torch.device(device)
        self.dist = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.dist else 0
        self.world = dist.get_world_size() if self.dist else 1
        if self.world > _C.PEER_MAX_WORLD:
            raise RuntimeError(f"PeerGroup: at most {_C.PEER_MAX_WORLD} ranks (one NVSwitch node)")
        self._bufs = []
        self.flags = self.alloc(_C.PEER_CHANNELS * _C.PEER_MAX_WORLD * 4)
        self.epoch = torch.zeros(_C.PEER_CHANNELS, dtype=torch.int32, device=self.device)
        self._bn_ctx = None

    def alloc(self, nbytes: int) -> PeerBuffer:
        import ctypes
        from . import _C
        nbytes = (int(nbytes) + 255) // 256 * 256
        with torch.cuda.device(self.device):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _C.check(_C.lib().vdetr_peer_alloc(nbytes, ctypes.byref(ptr), handle))
            handles = [None] * self.world
            if self.dist:
                dist.all_gather_object(handles, bytes(handle.raw))
            ptrs = (ctypes.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    ptrs[r] = ptr.value
                else:
                    q = ctypes.c_void_p()
                    _C.check(_C.lib().vdetr_peer_open(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(q)))
                    ptrs[r] = q.value
            local = torch.as_tensor(_RawCuda(ptr.value, nbytes), device=self.device)
        buf = PeerBuffer(local, ptrs, ptr.value)
        self._bufs.append(buf)
        if self.dist:
            dist.barrier()           # every rank has mapped every allocation before anyone uses it
        return buf

    def close(self):
        """Unmap the peers' allocations and free this rank's (after a barrier: nobody may still be using them).  Optional --
        the allocations otherwise live until the process exits -- and only legal once every tensor that views them is gone:
        the parameters, gradients and flat vectors of a FlatAdamW(peer=self) live in these buffers."""
        from . import _C
        if self._bn_ctx is not None:
            self.disable_sync_batchnorm()
        torch.cuda.synchronize(self.device)
        if self.dist:
            dist.barrier()
        with torch.cuda.device(self.device):
            for buf in self._bufs:
                for r in range(self.world):
                    if r != self.rank:
                        _C.check(_C.lib().vdetr_peer_close(buf.ptrs[r]))
            if self.dist:
                dist.barrier()
            for buf in self._bufs:
                buf.local = None
                _C.check(_C.lib().vdetr_peer_free(buf._raw))
        self._bufs = []

    def barrier(self, channel: int = 0):
        """Cross-GPU barrier on the current stream (a one-warp kernel; capturable in a CUDA graph)."""
        from . import _C
        with torch.cuda.device(self.device):
            _C.check(_C.lib().vdetr_peer_barrier(self.flags.ptrs, self.rank, self.world, _C.ptr(self.epoch), channel, _C.stream_ptr()))

    def check(self):
        """Raises if a barrier of this process ever timed out waiting for a peer."""
        import ctypes
        from . import _C
        e = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _C.check(_C.lib().vdetr_peer_error(ctypes.byref(e)))
        if e.value:
            raise RuntimeError("PeerGroup: a cross-GPU barrier timed out (a peer rank died or fell out of step)")

    def enable_sync_batchnorm(self, max_channels: int = 4096):
        """The reference converts every BatchNorm to SyncBatchNorm (main.py:512-514).  Here: from now on every training-mode
        BatchNorm+ReLU launch of this process (ops.bn_relu_train*) merges its statistics with the other ranks through peer
        memory -- one small kernel per exchange (csrc/peer.cu), no NCCL call.  Every rank must run the same layers."""
        from . import _C
        slots = self.alloc(_C.lib().vdetr_peer_bn_slot_floats(max_channels) * 4)
        ctx = _C.PeerCtx()
        for r in range(self.world):
            ctx.flags[r] = self.flags.ptrs[r]
            ctx.slots[r] = slots.ptrs[r]
        ctx.rank, ctx.world, ctx.epoch, ctx.cap = self.rank, self.world, _C.ptr(self.epoch), max_channels
        import ctypes
        _C.check(_C.lib().vdetr_bn_sync_set(ctypes.byref(ctx)))
        self._bn_ctx = ctx
        from . import ops
        ops.SYNC_BN_ACTIVE = True

    def disable_sync_batchnorm(self):
        from . import _C
        _C.check(_C.lib().vdetr_bn_sync_set(None))
        self._bn_ctx = None
        from . import ops
        ops.SYNC_BN_ACTIVE = False


def _peer_norm_bytes():
    from . import _C
    return int(_C.lib().vdetr_peer_norm_bytes())


class FlatAdamW:
    """torch.optim.AdamW semantics (the reference: optimizer.py:4-26, stepped after clip_grad_norm_ in engine.py:105-108) on
    FLAT buffers: parameters are re-homed as views of one fp32 buffer, gradients are the views of FlatGradAllReduce, the two
    moments are flat, and a step is ONE kernel launch (csrc/optim.cu) instead of ~22 multi-tensor launches over ~1000 small
    tensors.  lr, the step count and the gradient scale live on the device, so the step can sit inside a captured CUDA graph
    and still follow a learning-rate schedule (set_lr) or a clipped gradient norm.

    no_decay(name, param) -> True puts a parameter into the group without weight decay (--filter_biases_wd).

    peer = PeerGroup: the data-parallel form.  Parameters and gradients live in peer memory, every rank owns a contiguous
    shard of the flat vector, and step() is the fused exchange of csrc/peer.cu: reduce-scatter of the gradients by P2P loads,
    global gradient norm (max_norm > 0 clips like engine.py:105-106), AdamW on the shard with shard-sized moments, all-gather
    of the parameters by P2P stores -- no NCCL call, no separate all-reduce (FlatGradAllReduce.sync_ is not used)."""

    def __init__(self, named_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, no_decay=None, peer=None):
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
        self.peer = peer
        if peer is None:
            self.flat_p = torch.empty(self.n, dtype=torch.float32, device=ref.device)
        else:
            self._p_buf = peer.alloc(self.n * 4)
            self._g_buf = peer.alloc(self.n * 4)
            self.flat_p = self._p_buf.view(torch.float32, self.n)
        o = 0
        with torch.no_grad():
            for p in self.params:                      # parameters become views of the flat buffer (same values)
                v = self.flat_p[o:o + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
                o += p.numel()
        self.max_norm = 0.0                            # peer form: > 0 clips the global gradient norm inside step()
        if peer is None:
            self.grads = FlatGradAllReduce(self.params)    # same order: flat_g[i] is the gradient of flat_p[i]
            self.lo, self.hi = 0, self.n
        else:
            if peer.dist:
                dist.broadcast(self.flat_p, 0)         # every rank starts from rank 0's parameters (what DDP does)
            self.grads = FlatGradAllReduce(self.params, flat=self._g_buf.view(torch.float32, self.n))
            self.lo, self.hi = flat_shard_range(self.n, peer.rank, peer.world)
            self._norm_buf = peer.alloc(_peer_norm_bytes())
            self._reduced = torch.zeros(max(self.hi - self.lo, 4), dtype=torch.float32, device=ref.device)
            self.last_norm = torch.zeros(1, dtype=torch.float32, device=ref.device)
        # moments: the whole vector, or (peer form) this rank's shard only
        self.exp_avg = torch.zeros(max(self.hi - self.lo, 4), dtype=torch.float32, device=ref.device)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.lr = torch.tensor([lr], dtype=torch.float32, device=ref.device)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=ref.device)
        self.grad_scale = torch.ones(1, dtype=torch.float32, device=ref.device)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.world_scale = 1.0                         # 1 / world size when the all-reduce is a plain SUM

    def set_lr(self, lr: float):
        self.lr.fill_(lr)

    def clip_grad_norm_(self, max_norm: float):
        """torch.nn.utils.clip_grad_norm_ (engine.py:105-106) as a device-side scale of the flat gradient; returns the norm.
        Peer form: the norm of the summed gradient only exists inside step(); this sets max_norm for it and returns the
        device tensor step() writes the norm to."""
        if self.peer is not None:
            self.max_norm = float(max_norm)
            return self.last_norm
        norm = torch.linalg.vector_norm(self.grads.gather_()) * self.world_scale
        torch.clamp(max_norm / (norm + 1e-6), max=1.0, out=self.grad_scale[0])
        return norm

    @torch.no_grad()
    def step(self):
        from . import _C
        self.grads.gather_()
        self.step_t += 1.0
        if self.peer is not None:
            pg = self.peer
            with torch.cuda.device(self.flat_p.device):
                _C.check(_C.lib().vdetr_adamw_flat_peer(self._p_buf.ptrs, self._g_buf.ptrs, pg.flags.ptrs, self._norm_buf.ptrs, pg.rank,
                                                        pg.world, _C.ptr(pg.epoch), _C.ptr(self._reduced), _C.ptr(self.exp_avg),
                                                        _C.ptr(self.exp_avg_sq), self.n, self.n_decay, self.lo, self.hi, _C.ptr(self.lr),
                                                        _C.ptr(self.step_t), float(self.world_scale), float(self.max_norm),
                                                        _C.ptr(self.last_norm), float(self.betas[0]), float(self.betas[1]),
                                                        float(self.eps), float(self.weight_decay), _C.stream_ptr()))
            return
        with torch.cuda.device(self.flat_p.device):
            _C.check(_C.lib().vdetr_adamw_flat(_C.ptr(self.flat_p), _C.ptr(self.grads.flat), _C.ptr(self.exp_avg),
                                               _C.ptr(self.exp_avg_sq), self.n, self.n_decay, _C.ptr(self.lr), _C.ptr(self.step_t),
                                               _C.ptr(self.grad_scale), float(self.world_scale), float(self.betas[0]),
                                               float(self.betas[1]), float(self.eps), float(self.weight_decay), _C.stream_ptr()))

    def zero_grad(self, set_to_none: bool = True):
        self.grads.zero_()

    def _full_moment(self, shard):
        """Peer form: assemble a whole-vector moment from the shards of all ranks (checkpointing only)."""
        n = self.hi - self.lo
        if self.peer is None or not self.peer.dist:
            return shard[:n].clone()
        parts = [None] * self.peer.world
        dist.all_gather_object(parts, (self.lo, shard[:n].cpu()))
        full = torch.empty(self.n, dtype=torch.float32)
        for lo, t in parts:
            full[lo:lo + t.numel()] = t
        return full.to(shard.device)

    def state_dict(self):
        return {"exp_avg": self._full_moment(self.exp_avg), "exp_avg_sq": self._full_moment(self.exp_avg_sq),
                "step": self.step_t.clone(), "lr": self.lr.clone(), "names": list(self.names)}

    def load_state_dict(self, sd):
        if list(sd["names"]) != self.names:
            raise ValueError("FlatAdamW.load_state_dict: parameter order differs")
        n = self.hi - self.lo
        self.exp_avg[:n].copy_(sd["exp_avg"][self.lo:self.hi]); self.exp_avg_sq[:n].copy_(sd["exp_avg_sq"][self.lo:self.hi])
        self.step_t.copy_(sd["step"]); self.lr.copy_(sd["lr"])
