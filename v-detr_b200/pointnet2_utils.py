"""Drop-in for the operator surface of /root/reference/third_party/pointnet2/pointnet2_utils.py:48-288
(furthest_point_sample, gather_operation, ball_query, grouping_operation) on top of libvdetr_b200.

``_ext`` mirrors the pybind module of the reference (src/bindings.cpp:9-21): same function names, argument
order, dtypes, shapes, fresh output tensors, RuntimeError on CPU / non-contiguous / wrong-dtype input.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _C


class _Ext:
    """pointnet2._ext look-alike."""

    @staticmethod
    def furthest_point_sampling(points, nsamples):
        _C.require_cuda("points", points, torch.float32)
        B, N, _ = points.shape
        out = torch.zeros(B, nsamples, dtype=torch.int32, device=points.device)
        L = _C.lib()
        ws_bytes = L.vdetr_pn2_fps_workspace_bytes(B, N, nsamples)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device) if ws_bytes else None
        with torch.cuda.device(points.device):
            _C.check(L.vdetr_pn2_fps(_C.ptr(points), B, N, nsamples, _C.ptr(out), _C.ptr(ws), ws_bytes, _C.stream_ptr()))
        return out

    @staticmethod
    def furthest_point_sampling_ragged(points, offsets, nsamples, max_n):
        """points [total,3] f32, offsets [B+1] int32 (device), max_n = upper bound of the per-scene count (host int).
        -> int32 [B, nsamples] of scene-local indices.  No counterpart in the reference's _ext: it replaces the per-scene
        loop of B = 1 calls at models/model_vdetr.py:282-316."""
        _C.require_cuda("points", points, torch.float32)
        _C.require_cuda("offsets", offsets, torch.int32)
        if points.dim() != 2 or points.shape[1] != 3 or offsets.dim() != 1 or offsets.numel() < 2:
            raise RuntimeError("points must be [total,3] and offsets [B+1]")
        B, total = offsets.numel() - 1, points.shape[0]
        out = torch.zeros(B, nsamples, dtype=torch.int32, device=points.device)
        L = _C.lib()
        ws_bytes = L.vdetr_pn2_fps_ragged_workspace_bytes(B, total, int(max_n), nsamples)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=points.device) if ws_bytes else None
        with torch.cuda.device(points.device):
            _C.check(L.vdetr_pn2_fps_ragged(_C.ptr(points), _C.ptr(offsets), B, total, int(max_n), nsamples, _C.ptr(out),
                                            _C.ptr(ws), ws_bytes, _C.stream_ptr()))
        return out

    @staticmethod
    def gather_points(points, idx):
        _C.require_cuda("points", points, torch.float32)
        _C.require_cuda("idx", idx, torch.int32)
        B, C, N = points.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32, device=points.device)
        with torch.cuda.device(points.device):
            _C.check(_C.lib().vdetr_pn2_gather(_C.ptr(points), _C.ptr(idx), B, C, N, M, _C.ptr(out), _C.stream_ptr()))
        return out

    @staticmethod
    def gather_points_grad(grad_out, idx, n):
        _C.require_cuda("grad_out", grad_out, torch.float32)
        _C.require_cuda("idx", idx, torch.int32)
        B, C, M = grad_out.shape
        out = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _C.check(_C.lib().vdetr_pn2_gather_grad(_C.ptr(grad_out), _C.ptr(idx), B, C, n, M, _C.ptr(out), _C.stream_ptr()))
        return out

    @staticmethod
    def ball_query(new_xyz, xyz, radius, nsample):
        _C.require_cuda("new_xyz", new_xyz, torch.float32)
        _C.require_cuda("xyz", xyz, torch.float32)
        B, N, _ = xyz.shape
        M = new_xyz.shape[1]
        out = torch.empty(new_xyz.shape[0], M, nsample, dtype=torch.int32, device=xyz.device)
        with torch.cuda.device(xyz.device):
            _C.check(_C.lib().vdetr_pn2_ball_query(_C.ptr(new_xyz), _C.ptr(xyz), B, N, M, float(radius), nsample,
                                                   _C.ptr(out), _C.stream_ptr()))
        return out

    @staticmethod
    def group_points(points, idx):
        _C.require_cuda("points", points, torch.float32)
        _C.require_cuda("idx", idx, torch.int32)
        B, C, N = points.shape
        _, M, S = idx.shape
        out = torch.empty(B, C, M, S, dtype=torch.float32, device=points.device)
        with torch.cuda.device(points.device):
            _C.check(_C.lib().vdetr_pn2_group(_C.ptr(points), _C.ptr(idx), B, C, N, M, S, _C.ptr(out), _C.stream_ptr()))
        return out

    @staticmethod
    def group_points_grad(grad_out, idx, n):
        _C.require_cuda("grad_out", grad_out, torch.float32)
        _C.require_cuda("idx", idx, torch.int32)
        B, C, M, S = grad_out.shape
        out = torch.zeros(B, C, n, dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _C.check(_C.lib().vdetr_pn2_group_grad(_C.ptr(grad_out), _C.ptr(idx), B, C, n, M, S, _C.ptr(out), _C.stream_ptr()))
        return out


_ext = _Ext()


class FurthestPointSampling(Function):
    """pointnet2_utils.py:48-77 (indices are not differentiable)."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """pointnet2_utils.py:80-114."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(1), features.size(2))
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        return _ext.gather_points_grad(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply


class GroupingOperation(Function):
    """pointnet2_utils.py:206-254."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.for_backwards = (idx, features.size(2))
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, N = ctx.for_backwards
        return _ext.group_points_grad(grad_out.contiguous(), idx, N), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """pointnet2_utils.py:257-288 -- note the wrapper's argument order (radius, nsample, xyz, new_xyz)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply
