"""Autograd wrappers around the attention kernels of libvdetr_b200.

``rpe_attention`` is the fused core of GlobalShareCrossAttention.forward
(/root/reference/models/vdetr_transformer.py:708-753): Vertex-RPE bias + QK^T + softmax + PV, forward and
backward, without materialising the [B,H,nQ,nK] bias / attention tensors.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from . import _C

IMPL_TCGEN05 = 0     # product kernels (tcgen05 + TMA)
IMPL_SIMT = 1        # validation kernels


def default_impl() -> int:
    return int(os.environ.get("VDETR_B200_IMPL", IMPL_TCGEN05))


def save_bias_enabled() -> bool:
    """VDETR_B200_SAVE_BIAS=0 trades the 16 B / pair of saved bias for a recompute in the backward."""
    return os.environ.get("VDETR_B200_SAVE_BIAS", "1") != "0"


def _shape(q, k, tables, log_scale, max_value, rotate, has_bias):
    B, nQ, H, hd = q.shape
    nK, kvh = k.shape[1], k.shape[2]
    n = tables.shape[1] if has_bias else 0
    return _C.XattnShape(B, nQ, nK, H, hd, n, float(log_scale), float(max_value), int(rotate), kvh, int(has_bias))


def _validate(q, k, v, xyz, ref_pts, ref_angle, tables):
    """Shape contract of the kernels (they index raw pointers: a mismatch must never reach the device)."""
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise RuntimeError("q must be [B,nQ,H,hd], k / v must be [B,nK,kv_heads,hd]")
    B, nQ, H, hd = q.shape
    if (H, hd) != (4, 64):
        raise RuntimeError(f"the attention kernels are built for 4 heads x 64 channels, got {H} x {hd}")
    if k.shape != v.shape or k.shape[0] != B or k.shape[3] != hd or k.shape[2] not in (1, H):
        raise RuntimeError(f"k {tuple(k.shape)} / v {tuple(v.shape)} do not match q {tuple(q.shape)}")
    if tables is None:
        return
    nK = k.shape[1]
    if k.shape[2] != 1:
        raise RuntimeError("the Vertex-RPE bias requires one shared K/V head (kv_heads = 1)")
    if tables.dim() != 5 or tables.shape[0] != 8 or tables.shape[4] != H or not (tables.shape[1] == tables.shape[2] == tables.shape[3]):
        raise RuntimeError(f"tables must be [8,n,n,n,{H}], got {tuple(tables.shape)}")
    if xyz is None or tuple(xyz.shape) != (B, nK, 3):
        raise RuntimeError(f"xyz must be [{B},{nK},3]")
    if ref_pts is None or tuple(ref_pts.shape) != (B, nQ, 8, 3):
        raise RuntimeError(f"ref_pts must be [{B},{nQ},8,3]")
    if ref_angle is not None and tuple(ref_angle.shape) != (B, nQ):
        raise RuntimeError(f"ref_angle must be [{B},{nQ}]")


class _RpeAttention(Function):
    @staticmethod
    def forward(ctx, q, k, v, xyz, ref_pts, ref_angle, tables, log_scale, max_value, impl, impl_bwd, dropout_p, dropout_seed):
        has_bias = tables is not None
        _validate(q, k, v, xyz, ref_pts, ref_angle, tables)
        if dropout_p > 0.0:
            if impl != IMPL_TCGEN05 or impl_bwd != IMPL_TCGEN05:
                raise RuntimeError("attention dropout is implemented by the tcgen05 kernels only (impl 0)")
            _C.require_cuda("dropout_seed", dropout_seed, torch.int64)
        for name, t in (("q", q), ("k", k), ("v", v)):
            _C.require_cuda(name, t, torch.float32)
        if has_bias:
            for name, t in (("xyz", xyz), ("ref_pts", ref_pts), ("tables", tables)):
                _C.require_cuda(name, t, torch.float32)
            if ref_angle is not None:
                _C.require_cuda("ref_angle", ref_angle, torch.float32)
        s = _shape(q, k, tables, log_scale, max_value, ref_angle is not None, has_bias)
        out = torch.empty_like(q)
        lse = torch.empty(s.B, s.H, s.nQ, dtype=torch.float32, device=q.device)
        L = _C.lib()
        bias_save = None
        with torch.cuda.device(q.device):
            nbytes = L.vdetr_xattn_fwd_workspace_bytes(s, impl)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device) if nbytes else None
            # training: keep the per-pair bias (16 B / pair) so that the backward streams it instead of recomputing it
            if has_bias and impl == impl_bwd == IMPL_TCGEN05 and any(ctx.needs_input_grad) and save_bias_enabled():
                nsave = L.vdetr_xattn_bias_save_bytes(s, impl)
                if nsave:
                    bias_save = torch.empty(nsave, dtype=torch.uint8, device=q.device)
            _C.check(L.vdetr_xattn_fwd(s, _C.ptr(q), _C.ptr(k), _C.ptr(v), _C.ptr(xyz), _C.ptr(ref_pts), _C.ptr(ref_angle),
                                       _C.ptr(tables), _C.ptr(out), _C.ptr(lse), _C.ptr(bias_save), float(dropout_p),
                                       _C.ptr(dropout_seed), _C.ptr(ws), nbytes, impl, _C.stream_ptr()))
        ctx.save_for_backward(q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, bias_save, dropout_seed)
        ctx.meta = (log_scale, max_value, impl_bwd, has_bias, float(dropout_p))
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, xyz, ref_pts, ref_angle, tables, out, lse, bias_save, dropout_seed = ctx.saved_tensors
        log_scale, max_value, impl, has_bias, dropout_p = ctx.meta
        dout = dout.contiguous()
        s = _shape(q, k, tables, log_scale, max_value, ref_angle is not None, has_bias)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        dtab = torch.empty_like(tables) if has_bias else None
        L = _C.lib()
        with torch.cuda.device(q.device):
            nbytes = L.vdetr_xattn_bwd_workspace_bytes(s, impl, int(bias_save is not None))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device) if nbytes else None
            _C.check(L.vdetr_xattn_bwd(s, _C.ptr(q), _C.ptr(k), _C.ptr(v), _C.ptr(xyz), _C.ptr(ref_pts), _C.ptr(ref_angle),
                                       _C.ptr(tables), _C.ptr(out), _C.ptr(lse), _C.ptr(dout), _C.ptr(bias_save),
                                       float(dropout_p), _C.ptr(dropout_seed), _C.ptr(dq), _C.ptr(dk), _C.ptr(dv), _C.ptr(dtab),
                                       _C.ptr(ws), nbytes, impl, _C.stream_ptr()))
        return dq, dk, dv, None, None, None, dtab, None, None, None, None, None, None


def new_dropout_seed(device) -> torch.Tensor:
    """One 64-bit seed on `device`, drawn from torch's CUDA generator (so torch.manual_seed controls it and the draw is
    legal inside CUDA-graph capture: every replay advances the generator and re-seeds the attention dropout)."""
    return torch.randint(-(2 ** 62), 2 ** 62, (1,), dtype=torch.int64, device=device)


def rpe_attention(q, k, v, xyz=None, ref_pts=None, ref_angle=None, tables=None, log_scale=512.0, max_value=4.0,
                  impl=None, impl_bwd=None, dropout_p=0.0, dropout_seed=None):
    """softmax_k(q k^T + rpe(ref_pts, xyz, tables)) v, with optional dropout on the probabilities.

    q [B,nQ,H,hd] (pre-scaled by hd^-0.5), k/v [B,nK,kvh,hd] (kvh = 1: shared K/V head, kvh = H: per head),
    xyz [B,nK,3], ref_pts [B,nQ,8,3], ref_angle [B,nQ] or None, tables [8,n,n,n,H] or None (no bias).
    dropout_p > 0: nn.Dropout(p) on the attention probabilities (vdetr_transformer.py:751-752) inside the kernels;
    dropout_seed = int64 CUDA tensor [1] (default: a fresh draw from torch's generator).
    Returns [B,nQ,H,hd].  Gradients: q, k, v, tables (xyz / ref_pts are detached in the reference).
    """
    dropout_p = float(dropout_p)
    if not 0.0 <= dropout_p < 1.0:
        raise RuntimeError(f"dropout_p must be in [0, 1), got {dropout_p}")
    if dropout_p > 0.0 and dropout_seed is None:
        dropout_seed = new_dropout_seed(q.device)
    if dropout_p == 0.0:
        dropout_seed = None
    impl = default_impl() if impl is None else impl
    impl_bwd = int(os.environ.get("VDETR_B200_IMPL_BWD", impl)) if impl_bwd is None else impl_bwd
    return _RpeAttention.apply(q.contiguous(), k.contiguous(), v.contiguous(),
                               None if xyz is None else xyz.contiguous(),
                               None if ref_pts is None else ref_pts.contiguous(),
                               None if ref_angle is None else ref_angle.contiguous(),
                               None if tables is None else tables.contiguous(), log_scale, max_value, impl, impl_bwd,
                               dropout_p, dropout_seed)


def _validate_bias_args(xyz, ref_pts, ref_angle, tables):
    if xyz.dim() != 3 or xyz.shape[2] != 3:
        raise RuntimeError("xyz must be [B,nK,3]")
    B, nK = xyz.shape[:2]
    if ref_pts.dim() != 4 or ref_pts.shape[0] != B or tuple(ref_pts.shape[2:]) != (8, 3):
        raise RuntimeError(f"ref_pts must be [{B},nQ,8,3], got {tuple(ref_pts.shape)}")
    nQ = ref_pts.shape[1]
    if tables.dim() != 5 or tables.shape[0] != 8 or tables.shape[4] != 4 or not (tables.shape[1] == tables.shape[2] == tables.shape[3]):
        raise RuntimeError(f"tables must be [8,n,n,n,4] (4 heads), got {tuple(tables.shape)}")
    if ref_angle is not None and tuple(ref_angle.shape) != (B, nQ):
        raise RuntimeError(f"ref_angle must be [{B},{nQ}]")
    return B, nK, nQ


def rpe_bias(xyz, ref_pts, tables, ref_angle=None, log_scale=512.0, max_value=4.0):
    """Materialise rpe [B,H,nQ,nK] (debug / return_attn_weights path; vdetr_transformer.py:708-731)."""
    for name, t in (("xyz", xyz), ("ref_pts", ref_pts), ("tables", tables)):
        _C.require_cuda(name, t, torch.float32)
    B, nK, nQ = _validate_bias_args(xyz, ref_pts, ref_angle, tables)
    s = _C.XattnShape(B, nQ, nK, 4, 64, tables.shape[1], float(log_scale), float(max_value), int(ref_angle is not None), 1, 1)
    out = torch.empty(B, 4, nQ, nK, dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _C.check(_C.lib().vdetr_rpe_bias(s, _C.ptr(xyz), _C.ptr(ref_pts), _C.ptr(ref_angle), _C.ptr(tables), _C.ptr(out),
                                         _C.stream_ptr()))
    return out


def rpe_bias_grad_tables(xyz, ref_pts, ref_angle, tables, dbias, log_scale=512.0, max_value=4.0):
    """dTables [8,n,n,n,H] from a dense d(bias) [B,H,nQ,nK] (adjoint of ``rpe_bias``)."""
    B, nK, nQ = _validate_bias_args(xyz, ref_pts, ref_angle, tables)
    if tuple(dbias.shape) != (B, 4, nQ, nK):
        raise RuntimeError(f"dbias must be [{B},4,{nQ},{nK}], got {tuple(dbias.shape)}")
    s = _C.XattnShape(B, nQ, nK, 4, 64, tables.shape[1], float(log_scale), float(max_value), int(ref_angle is not None), 1, 1)
    ds4 = dbias.permute(0, 2, 3, 1).contiguous()
    out = torch.empty_like(tables)
    L = _C.lib()
    with torch.cuda.device(xyz.device):
        nbytes = L.vdetr_rpe_dtables_workspace_bytes(s)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device)
        _C.check(L.vdetr_rpe_dtables(s, _C.ptr(xyz.contiguous()), _C.ptr(ref_pts.contiguous()),
                                     _C.ptr(None if ref_angle is None else ref_angle.contiguous()), _C.ptr(ds4), _C.ptr(out),
                                     _C.ptr(ws), nbytes, _C.stream_ptr()))
    return out


def _reduce_ws(cols, device):
    """Scratch of the deterministic cross-CTA column reductions (partial sums + ticket); never read by the caller."""
    return torch.empty(_C.lib().vdetr_reduce_workspace_floats(cols), dtype=torch.float32, device=device)


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        cols = x.shape[-1]
        x2 = x.reshape(-1, cols)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _C.check(_C.lib().vdetr_layernorm_fwd(_C.ptr(x2), _C.ptr(weight), _C.ptr(bias), rows, cols, float(eps), _C.ptr(y),
                                                  _C.ptr(mean), _C.ptr(rstd), _C.stream_ptr()))
        ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        rows, cols = x2.shape
        dy2 = dy.contiguous().view(rows, cols)
        dx = torch.empty_like(x2)
        dw = torch.empty_like(weight)
        db = torch.empty_like(weight)
        with torch.cuda.device(x2.device):
            ws = _reduce_ws(cols, x2.device)
            _C.check(_C.lib().vdetr_layernorm_bwd(_C.ptr(dy2), _C.ptr(x2), _C.ptr(mean), _C.ptr(rstd), _C.ptr(weight), rows, cols,
                                                  _C.ptr(dx), _C.ptr(dw), _C.ptr(db), _C.ptr(ws), _C.stream_ptr()))
        return dx.view(ctx.shape), dw, db, None


def layer_norm_supported(x, weight, bias) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and weight is not None and bias is not None and weight.dim() == 1
            and weight.dtype == torch.float32 and x.shape[-1] == weight.shape[0] and x.shape[-1] in (128, 256, 384, 512))


# True while the library-wide SyncBatchNorm switch is on (parallel.PeerGroup.enable_sync_batchnorm): every training BatchNorm
# launch then contains a cross-GPU exchange on ONE flag channel, so BatchNorm launches must stay on one stream, in the same
# order on every rank (vdetr_transformer._run_heads does not fork streams then).
SYNC_BN_ACTIVE = False


def batch_first_backed(x) -> bool:
    """x is a [L, B, D] (sequence-first, the reference's convention) VIEW of contiguous batch-first memory [B, L, D].
    The decoder keeps its token activations that way: per-token ops (linear, layer norm, dropout, residual adds) run on the
    memory order and hand on the same kind of view, and the attention kernels -- which want [B, L, H, hd] -- get their
    operands without the transposing copies a sequence-first memory order costs (4 per attention call and direction)."""
    return x.dim() == 3 and not x.is_contiguous() and x.transpose(0, 1).is_contiguous()


def layer_norm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm over the last dimension (fp32, CUDA) through the library's warp-per-row kernels."""
    if batch_first_backed(x):
        return layer_norm(x.transpose(0, 1), weight, bias, eps).transpose(0, 1)
    return _LayerNorm.apply(x.contiguous(), weight.contiguous(), bias.contiguous(), eps)


class _BnReluTrain(Function):
    """relu(BatchNorm(x)) for `groups` independent BatchNorm layers of `cols` channels in one pair of kernels.
    layout "cl": x [rows, groups * cols] (channels last); layout "gm": x [groups, rows, cols] (group major)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, groups, layout, dropout_p, dropout_seed):
        if layout == "cl":
            rows, cols = x.shape[0], x.shape[1] // groups
            gstride, rstride = cols, x.shape[1]
        else:
            rows, cols = x.shape[1], x.shape[2]
            gstride, rstride = rows * cols, cols
        y = torch.empty_like(x)
        mean = torch.empty(groups * cols, dtype=torch.float32, device=x.device)
        rstd = torch.empty(groups * cols, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            ws = _reduce_ws(groups * cols, x.device)
            _C.check(_C.lib().vdetr_bn_relu_train_fwd(_C.ptr(x), _C.ptr(weight), _C.ptr(bias), rows, cols, groups, gstride, rstride,
                                                      float(eps), float(momentum), float(dropout_p), _C.ptr(dropout_seed), _C.ptr(y),
                                                      _C.ptr(mean), _C.ptr(rstd), _C.ptr(running_mean), _C.ptr(running_var),
                                                      _C.ptr(ws), _C.stream_ptr()))
        ctx.save_for_backward(x, y, weight, mean, rstd)
        ctx.geom = (rows, cols, groups, gstride, rstride, float(dropout_p))
        ctx.mark_non_differentiable(*[t for t in (running_mean, running_var) if t is not None])
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, rstd = ctx.saved_tensors
        rows, cols, groups, gstride, rstride, dropout_p = ctx.geom
        dy = dy.contiguous()
        dx, dw, db = torch.empty_like(x), torch.empty_like(weight), torch.empty_like(weight)
        with torch.cuda.device(x.device):
            ws = _reduce_ws(groups * cols, x.device)
            _C.check(_C.lib().vdetr_bn_relu_train_bwd(_C.ptr(dy), _C.ptr(y), _C.ptr(x), _C.ptr(mean), _C.ptr(rstd), _C.ptr(weight),
                                                      rows, cols, groups, gstride, rstride, dropout_p, _C.ptr(dx), _C.ptr(dw),
                                                      _C.ptr(db), _C.ptr(ws), _C.stream_ptr()))
        return dx, dw, db, None, None, None, None, None, None, None, None


def bn_relu_train_supported(x, bn) -> bool:
    return (bn.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[0] > 1 and bn.affine
            and bn.momentum is not None and x.shape[1] in (128, 256, 384, 512) and bn.weight.dtype == torch.float32)


def _drop_args(dropout_p, device):
    dropout_p = float(dropout_p)
    return (dropout_p, new_dropout_seed(device)) if dropout_p > 0.0 else (0.0, None)


def bn_relu_train(x, bn, dropout_p=0.0):
    """dropout(relu(BatchNorm1d(x))) for a training-mode nn.BatchNorm1d `bn` on token-major x [T, C] (fp32, CUDA): batch
    statistics, running-statistics update, ReLU and (dropout_p > 0) the Dropout that follows in the reference's stacks, in two
    kernels forward / two backward (csrc/batchnorm.cu)."""
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _BnReluTrain.apply(x.contiguous(), bn.weight, bn.bias, rm, rv, bn.eps, bn.momentum, 1, "cl", *_drop_args(dropout_p, x.device))


def bn_relu_train_group_supported(x, bns, layout) -> bool:
    b0 = bns[0]
    cols = b0.num_features
    same = all(type(b) is torch.nn.BatchNorm1d and b.training and b.affine and b.num_features == cols and b.eps == b0.eps
               and b.momentum == b0.momentum and b.momentum is not None and b.track_running_stats == b0.track_running_stats
               for b in bns)
    shape_ok = (x.dim() == 2 and x.shape[1] == cols * len(bns) and x.shape[0] > 1) if layout == "cl" else \
        (x.dim() == 3 and x.shape[0] == len(bns) and x.shape[2] == cols and x.shape[1] > 1)
    return same and shape_ok and x.is_cuda and x.dtype == torch.float32 and cols in (128, 256, 384, 512) and len(bns) <= 32


def bn_relu_train_group(x, bns, layout, dropout_p=0.0):
    """relu(BatchNorm1d_g(x_g)) for several training-mode nn.BatchNorm1d layers of equal width at once: x is
    [T, G * C] (layout "cl", the output of one GEMM with the G weight matrices concatenated) or [G, T, C] ("gm", the
    output of a batched GEMM).  Parameters are concatenated for the call (autograd splits the gradients back); the
    running statistics are updated through one temporary and copied back to the modules."""
    weight = torch.cat([b.weight for b in bns])
    bias = torch.cat([b.bias for b in bns])
    track = bns[0].track_running_stats
    rm = torch.cat([b.running_mean for b in bns]) if track else None
    rv = torch.cat([b.running_var for b in bns]) if track else None
    y = _BnReluTrain.apply(x.contiguous(), weight, bias, rm, rv, bns[0].eps, bns[0].momentum, len(bns), layout,
                           *_drop_args(dropout_p, x.device))
    if track:
        with torch.no_grad():
            c = bns[0].num_features
            torch._foreach_copy_([b.running_mean for b in bns] + [b.running_var for b in bns],
                                 list(rm.split(c)) + list(rv.split(c)))
            nbt = [b.num_batches_tracked for b in bns if b.num_batches_tracked is not None]
            if nbt:
                torch._foreach_add_(nbt, 1)
    return y


def _wgrad(dy2, x2):
    """dW = dy2^T x2 for token-major operands ([T, out], [T, in]).  The contraction runs over the T = 8192 ... 32768 tokens and
    the result is at most 256 x 256: a plain GEMM call tiles only the output (8 CTAs on 148 SMs, 16-140 us per call, 300
    calls per step).  Split-K as a batched GEMM over 512-token slices + one fixed-order sum spreads the same work over
    T / 512 x 8 CTAs and stays bit-reproducible."""
    T = dy2.shape[0]
    S = T // 512
    if x2.shape[1] % 4 and x2.shape[1] < 64:       # 3- / 6-channel inputs: keep cuBLAS on its aligned kernels
        pad = -x2.shape[1] % 4
        return _wgrad(dy2, torch.nn.functional.pad(x2, (0, pad)))[:, :x2.shape[1]]
    if S >= 4 and T % 512 == 0 and dy2.shape[1] * x2.shape[1] <= 256 * 512 and dy2.is_contiguous() and x2.is_contiguous():
        S = min(S, 64)
        while T % S:
            S -= 1
        return torch.bmm(dy2.view(S, T // S, -1).transpose(1, 2), x2.view(S, T // S, -1)).sum(0)
    return dy2.t() @ x2


class _TokenLinear(Function):
    """F.linear on [..., in] features whose bias gradient is the library's column-sum kernel (the GEMMs stay cuBLAS; the
    weight gradient is a split-K batched GEMM, see _wgrad)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        x2 = x.reshape(-1, x.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (dy2 @ weight).view(x.shape)
        if ctx.needs_input_grad[1]:
            dw = _wgrad(dy2, x2 if x2.is_contiguous() else x2.contiguous())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(dy2.shape[1], dtype=dy2.dtype, device=dy2.device)
            with torch.cuda.device(dy2.device):
                ws = torch.empty(_C.lib().vdetr_colsum_workspace_floats(dy2.shape[1]), dtype=torch.float32, device=dy2.device)
                _C.check(_C.lib().vdetr_colsum(_C.ptr(dy2), dy2.shape[0], dy2.shape[1], _C.ptr(db), _C.ptr(ws), _C.stream_ptr()))
        return dx, dw, db


class _BatchedTokenLinear(Function):
    """y[g] = x[g] @ W[g]^T (+ b[g]) for G independent layers on token-major activations: x [G, T, in] (any strides a batched
    GEMM takes), W [G, out, in], b [G, out] or None -> [G, T, out].  What torch.bmm / baddbmm with W.transpose(1, 2) compute, but
    the backward produces dW directly in the parameters' [G, out, in] layout (autograd's BmmBackward returns the transpose:
    one strided copy per head in AccumulateGrad) and as a split-K batched GEMM over 512-token slices + one fixed-order sum
    (the contraction runs over T = 8192 ... 32768 tokens for a 256 x 256 result: G x 4 output tiles otherwise)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if bias is None:
            return torch.bmm(x, weight.transpose(1, 2))
        return torch.baddbmm(bias.unsqueeze(1), x, weight.transpose(1, 2))

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        G, T, O = dy.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.bmm(dy, weight)
        if ctx.needs_input_grad[1]:
            S = min(T // 512, 64)
            while S > 1 and T % S:
                S -= 1
            if S >= 4 and dy.is_contiguous() and x.is_contiguous():
                part = torch.bmm(dy.view(G * S, T // S, O).transpose(1, 2), x.view(G * S, T // S, -1))      # [G*S, out, in]
                dw = part.view(G, S, O, -1).sum(1)
            else:
                dw = torch.bmm(dy.transpose(1, 2), x)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(1)
        return dx, dw, db


def batched_linear(x, weight, bias=None):
    """y[g] = x[g] @ weight[g]^T + bias[g] for x [G, T, in], weight [G, out, in], bias [G, out] (see _BatchedTokenLinear)."""
    if x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        return _BatchedTokenLinear.apply(x, weight, bias)
    if bias is None:
        return torch.bmm(x, weight.transpose(1, 2))
    return torch.baddbmm(bias.unsqueeze(1), x, weight.transpose(1, 2))


def linear(x, weight, bias=None):
    """torch.nn.functional.linear; on fp32 CUDA tensors that require grad the backward uses the library's column-sum
    kernel for the bias gradient."""
    if batch_first_backed(x):
        return linear(x.transpose(0, 1), weight, bias).transpose(0, 1)
    if (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and torch.is_grad_enabled()
            and (x.requires_grad or weight.requires_grad) and x.shape[-1] == weight.shape[1]):
        return _TokenLinear.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


class _BoxDecode(Function):
    """models/vdetr_transformer.py:244-333 for angle == 0 in one kernel per direction (csrc/boxdecode.cu)."""

    @staticmethod
    def forward(ctx, center_reg, size_reg, pre_cn, pre_sn, lo, hi):
        B, nQ, _ = center_reg.shape
        mk = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=center_reg.device)      # noqa: E731
        center, center_norm, size, size_norm, pre_center, pre_size = (mk(B, nQ, 3) for _ in range(6))
        corners, ref_lidar = mk(B, nQ, 8, 3), mk(B, nQ, 8, 3)
        with torch.cuda.device(center_reg.device):
            _C.check(_C.lib().vdetr_box_decode_fwd(_C.ptr(center_reg), _C.ptr(size_reg), _C.ptr(pre_cn), _C.ptr(pre_sn), _C.ptr(lo),
                                                   _C.ptr(hi), B, nQ, _C.ptr(center), _C.ptr(center_norm), _C.ptr(size),
                                                   _C.ptr(size_norm), _C.ptr(pre_center), _C.ptr(pre_size), _C.ptr(corners),
                                                   _C.ptr(ref_lidar), _C.stream_ptr()))
        ctx.save_for_backward(size, pre_size, lo, hi)
        ctx.mark_non_differentiable(pre_center, pre_size, ref_lidar)
        return center, center_norm, size, size_norm, corners, pre_center, pre_size, ref_lidar

    @staticmethod
    def backward(ctx, g_center, g_center_norm, g_size, g_size_norm, g_corners, _g5, _g6, _g7):
        size, pre_size, lo, hi = ctx.saved_tensors
        B, nQ, _ = size.shape
        c = lambda t: None if t is None else t.contiguous()      # noqa: E731
        g_center, g_center_norm, g_size, g_size_norm, g_corners = (c(t) for t in (g_center, g_center_norm, g_size, g_size_norm, g_corners))
        d_center_reg, d_size_reg = torch.empty_like(size), torch.empty_like(size)
        with torch.cuda.device(size.device):
            _C.check(_C.lib().vdetr_box_decode_bwd(_C.ptr(size), _C.ptr(pre_size), _C.ptr(lo), _C.ptr(hi), _C.ptr(g_center),
                                                   _C.ptr(g_center_norm), _C.ptr(g_size), _C.ptr(g_size_norm), _C.ptr(g_corners), B, nQ,
                                                   _C.ptr(d_center_reg), _C.ptr(d_size_reg), _C.stream_ptr()))
        return d_center_reg, d_size_reg, None, None, None, None


def box_decode_supported(center_reg, size_reg, pre_cn, pre_sn, lo, hi) -> bool:
    ts = (center_reg, size_reg, pre_cn, pre_sn, lo, hi)
    return all(t.is_cuda and t.dtype == torch.float32 for t in ts) and center_reg.dim() == 3 and center_reg.shape[-1] == 3 \
        and size_reg.shape == center_reg.shape and pre_cn.shape == center_reg.shape and pre_sn.shape == center_reg.shape \
        and tuple(lo.shape) == (center_reg.shape[0], 3) and tuple(hi.shape) == (center_reg.shape[0], 3)


def box_decode(center_reg, size_reg, pre_center_normalized, pre_size_normalized, dims_min, dims_max):
    """-> center, center_normalized, size, size_normalized, corners (camera frame), pre_center, pre_size, ref_lidar.
    Gradients flow to center_reg / size_reg only (the proposals and the scene extent are detached in the reference)."""
    return _BoxDecode.apply(center_reg.contiguous(), size_reg.contiguous(), pre_center_normalized.detach().contiguous(),
                            pre_size_normalized.detach().contiguous(), dims_min.detach().contiguous(), dims_max.detach().contiguous())
