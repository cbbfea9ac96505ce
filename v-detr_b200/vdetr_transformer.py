"""Drop-in for /root/reference/models/vdetr_transformer.py on top of the sm_100a kernels.

Same public classes, constructor arguments, ``forward`` signatures and state_dict keys:
``TransformerDecoder`` (:105-452), ``GlobalDecoderLayer`` (:455-582), ``FFNLayer`` (:585-606),
``ShareSelfAttention`` (:609-653), ``GlobalShareCrossAttention`` (:656-758), ``BoxProcessor`` (:20-90),
``convert_corners_camera2lidar`` (:98-102).

What is different underneath:
  * cross attention never materialises the [B,4,nQ,nK] bias or probability tensors: the eight vertex tables
    (cpb_mlps evaluated on the 10^3 lattice, stays PyTorch so autograd reaches the MLP weights through
    dTables) are handed to ``ops.rpe_attention`` -- one fused kernel forward, one backward;
  * decoder self attention (``nn.MultiheadAttention`` in the reference, :468) is ``MultiheadSelfAttention``:
    identical parameter names, the same fused kernel without bias;
  * key tokens are visited in Morton order (attention is permutation invariant over keys) so that the 32 keys a
    warp handles for one query fall into the same table cells and the table reads broadcast.
There is no CPU path: modules raise if their inputs are not CUDA tensors.
"""
from __future__ import annotations

import copy
import math
import os
from functools import partial
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import ops
from .helpers import ACTIVATION_DICT, NORM_DICT, WEIGHT_INIT_DICT, GenericMLP, Linear, PositionEmbeddingLearned, get_clones


# ------------------------------------------------------------------------------------------------- geometry
def shift_scale_points(pred_xyz, src_range, dst_range=None):
    """utils/pc_util.py:38-66 (only the default dst_range=[0,1] use of the decoder)."""
    lo, hi = src_range
    if dst_range is None:
        return (pred_xyz - lo[:, None, :]) / (hi[:, None, :] - lo[:, None, :])
    dlo, dhi = dst_range
    return (pred_xyz - lo[:, None, :]) * (dhi - dlo)[:, None, :] / (hi - lo)[:, None, :] + dlo[:, None, :]


def scale_points(pred_xyz, mult_factor):
    """utils/pc_util.py:69-73."""
    return pred_xyz * mult_factor[:, None, :]


_CORNER_SIGNS = {}


def _corner_signs(like):
    """Per-device constant sign vectors (cached: creating them from Python lists is a host->device copy, which is
    not allowed while a CUDA graph is being captured)."""
    key = (like.device, like.dtype)
    if key not in _CORNER_SIGNS:
        _CORNER_SIGNS[key] = tuple(torch.tensor(v, dtype=like.dtype, device=like.device) for v in (
            [1., 1., -1., -1., 1., 1., -1., -1.], [1., 1., 1., 1., -1., -1., -1., -1.], [1., -1., -1., 1., 1., -1., -1., 1.]))
    return _CORNER_SIGNS[key]


def box_corners_camera(center, size, angle):
    """dataset_config.box_parametrization_to_corners for ScanNet (datasets/scannet.py:168-171 ->
    utils/box_util.py:294-358): 8 corners in the camera frame, reference vertex order."""
    cam = torch.stack((center[..., 0], -center[..., 2], center[..., 1]), dim=-1)
    hl, hw, hh = size[..., 0:1] * 0.5, size[..., 1:2] * 0.5, size[..., 2:3] * 0.5
    sx, sy, sz = _corner_signs(size)
    lx, ly, lz = hl * sx, hh * sy, hw * sz
    c, s = torch.cos(angle).unsqueeze(-1), torch.sin(angle).unsqueeze(-1)
    # local @ roty(angle)^T written out (utils/box_util.py:304-317,352-356): a [.,8,3]x[.,3,3] batched matmul per box
    # costs a full GEMM launch per call and would run in TF32 when the caller allows it
    return torch.stack((c * lx + s * lz, ly, c * lz - s * lx), dim=-1) + cam.unsqueeze(-2)


class ScanNetBoxConfig:
    """The part of ScannetDatasetConfig the decoder reads (datasets/scannet.py:38-41,168-171).  The reference's
    own config object can be passed instead; only these attributes are used."""

    # mean box size (m) per ScanNet class: the data constants of datasets/scannet.py:72-91, used by ModelVDETR to
    # pick the proposal anchors from the per-point class prediction (models/model_vdetr.py:347-353)
    MEAN_SIZE = ((0.76966726, 0.81160211, 0.92573741), (1.876858, 1.84255952, 1.19315654), (0.61327999, 0.61486087, 0.71827014),
                 (1.39550063, 1.51215451, 0.83443565), (0.97949596, 1.06751485, 0.63296875), (0.53166301, 0.59555772, 1.75001483),
                 (0.96247056, 0.72462326, 1.14818682), (0.83221924, 1.04909355, 1.68756634), (0.21132214, 0.4206159, 0.53728459),
                 (1.44400728, 1.89708334, 0.26985747), (1.02942616, 1.40407966, 0.87554322), (1.37664116, 0.65521793, 1.68131292),
                 (0.66508189, 0.71111926, 1.29885307), (0.41999174, 0.37906947, 1.75139715), (0.59359559, 0.59124924, 0.73919014),
                 (0.50867595, 0.50656087, 0.30136236), (1.15115265, 1.0546296, 0.49706794), (0.47535286, 0.49249493, 0.58021168))

    def __init__(self, num_semcls=18, num_angle_bin=1):
        self.num_semcls = num_semcls
        self.num_angle_bin = num_angle_bin
        self.mean_size_arr = np.array(self.MEAN_SIZE)
        self.mean_size_arr_hard_anchor = np.ones((18, 3))

    @staticmethod
    def box_parametrization_to_corners(center, size, angle):
        return box_corners_camera(center, size, angle)


def convert_corners_camera2lidar(corners_camera):
    """(x_c, y_c, z_c) -> (x_c, z_c, -y_c); in place like the reference (:98-102)."""
    y = corners_camera[..., 1].clone()
    corners_camera[..., 1] = corners_camera[..., 2]
    corners_camera[..., 2] = -y
    return corners_camera


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def roty_batch_tensor(t):
    c, s = torch.cos(t), torch.sin(t)
    out = torch.zeros(tuple(t.shape) + (3, 3), dtype=torch.float32, device=t.device)
    out[..., 0, 0] = c; out[..., 0, 2] = s; out[..., 1, 1] = 1; out[..., 2, 0] = -s; out[..., 2, 2] = c
    return out


def rotz_batch_tensor(t):
    c, s = torch.cos(t), torch.sin(t)
    out = torch.zeros(tuple(t.shape) + (3, 3), dtype=torch.float32, device=t.device)
    out[..., 0, 0] = c; out[..., 0, 1] = -s; out[..., 1, 0] = s; out[..., 1, 1] = c; out[..., 2, 2] = 1
    return out


def morton_order(xyz: Tensor) -> Tensor:
    """Permutation that sorts the key tokens of every scene along a 30-bit Morton curve.  xyz [B,N,3] -> [B,N]."""
    lo = xyz.amin(dim=1, keepdim=True)
    span = (xyz.amax(dim=1, keepdim=True) - lo).clamp_min(1e-6)
    g = ((xyz - lo) / span * 1023.0).long().clamp_(0, 1023)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(g[..., 0]) | (spread(g[..., 1]) << 1) | (spread(g[..., 2]) << 2)
    return torch.argsort(code, dim=1)


_HEAD_STREAMS = {}


def _head_streams(device, n):
    """Side streams for the independent box heads of a level, cached per device at module level (stream objects in
    module state would break copy.deepcopy / torch.save of the model)."""
    key = (device.type, device.index)
    if key not in _HEAD_STREAMS or len(_HEAD_STREAMS[key]) < n:
        _HEAD_STREAMS[key] = [torch.cuda.Stream(device) for _ in range(n)]
    return _HEAD_STREAMS[key][:n]


class _PadStack(torch.autograd.Function):
    """torch.stack([F.pad(t, (..., 0, omax - t.shape[0])) for t in ts]) for tensors that differ in their first dimension:
    one zero fill + one copy per tensor forward (F.pad + stack: two launches per tensor + one), slices of the incoming gradient
    backward (no kernel)."""

    @staticmethod
    def forward(ctx, omax, *ts):
        ctx.sizes = [t.shape[0] for t in ts]
        out = ts[0].new_zeros((len(ts), omax) + tuple(ts[0].shape[1:]))
        for g, t in enumerate(ts):
            out[g, :t.shape[0]].copy_(t)
        return out

    @staticmethod
    def backward(ctx, go):
        return (None,) + tuple(go[g, :n] for g, n in enumerate(ctx.sizes))


class _SplitHeadOutputs(torch.autograd.Function):
    """o [G, T = nQ*B, omax] (token t = q*B + b) -> G tensors [B, nQ, outs[g]].  The backward writes the G gradients into one
    zero-filled [G, T, omax] buffer with one strided copy each; the equivalent chain of select / slice / view / transpose nodes
    costs ~5 launches per head and direction (45 head outputs per decoder pass)."""

    @staticmethod
    def forward(ctx, o, nQ, B, outs, batch_first=False):
        # batch_first: token t = b*nQ + q (the features came as a [nQ, B, C] view of batch-first memory)
        ctx.meta = (o.shape, nQ, B, tuple(outs), batch_first)
        if batch_first:
            return tuple(o[g, :, :w].reshape(B, nQ, w) for g, w in enumerate(outs))
        return tuple(o[g, :, :w].reshape(nQ, B, w).transpose(0, 1).contiguous() for g, w in enumerate(outs))

    @staticmethod
    def backward(ctx, *grads):
        shape, nQ, B, outs, batch_first = ctx.meta
        ref = next(g for g in grads if g is not None)
        go = ref.new_zeros(shape)
        for g, w in enumerate(outs):
            if grads[g] is not None:
                if batch_first:
                    go[g, :, :w].view(B, nQ, w).copy_(grads[g])
                else:
                    go[g, :, :w].view(nQ, B, w).copy_(grads[g].transpose(0, 1))
        return go, None, None, None, None


class BoxProcessor(object):
    """Turns MLP head outputs into boxes (:20-90)."""

    def __init__(self, dataset_config, cls_loss="celoss"):
        self.dataset_config = dataset_config
        self.cls_loss = cls_loss

    def compute_predicted_center(self, center_offset, query_xyz, point_cloud_dims):
        center_unnormalized = query_xyz + center_offset
        return shift_scale_points(center_unnormalized, src_range=point_cloud_dims), center_unnormalized

    def compute_predicted_class_size(self, size_normalized_offset, logits):
        class_idx = logits.sigmoid().max(dim=-1)[1]
        per_class = torch.tensor(self.dataset_config.mean_size_arr, device=logits.device).float()[class_idx]
        return per_class + size_normalized_offset * per_class, size_normalized_offset * per_class

    def compute_predicted_size(self, size_normalized, point_cloud_dims):
        scene = torch.clamp(point_cloud_dims[1] - point_cloud_dims[0], min=1e-1)
        return scale_points(size_normalized, mult_factor=scene)

    def compute_predicted_angle(self, angle_logits, angle_residual, zero_angle=False):
        nbin = angle_logits.shape[-1]
        if nbin == 1 or zero_angle:
            # keep the heads in the autograd graph (DDP), value is identically zero (:49-59)
            if nbin == 1:
                angle = (angle_logits * 0 + angle_residual * 0).squeeze(-1).clamp(min=0)
            else:
                angle = (angle_logits.sum(-1) * 0 + angle_residual.sum(-1) * 0).squeeze(-1).clamp(min=0)
            return angle, angle
        per_bin = 2 * np.pi / self.dataset_config.num_angle_bin
        prob, cls = F.softmax(angle_logits, dim=-1).max(dim=-1)
        cls = cls.detach()
        angle = per_bin * cls + angle_residual.gather(2, cls.unsqueeze(-1)).squeeze(-1)
        angle = torch.where(angle > np.pi, angle - 2 * np.pi, angle)
        return angle, prob

    def compute_objectness_and_cls_prob(self, cls_logits):
        if self.cls_loss.split("_")[0] == "focalloss":
            return cls_logits, cls_logits.sigmoid().max(dim=-1)[0]
        assert cls_logits.shape[-1] == self.dataset_config.num_semcls + 1
        prob = F.softmax(cls_logits, dim=-1)
        return prob[..., :-1], 1 - prob[..., -1]

    def box_parametrization_to_corners(self, box_center_unnorm, box_size_unnorm, box_angle):
        return self.dataset_config.box_parametrization_to_corners(box_center_unnorm, box_size_unnorm, box_angle)


# ------------------------------------------------------------------------------------------------- attention
def _dense_attention(q, k, v, bias, drop, mask=None):
    """Materialising attention used only for return_attn_weights / attn_mask (the fused kernels return no [B,H,nQ,nK]
    tensor and take no mask).  q [B,nQ,H,hd], k/v [B,nK,kvh,hd], bias [B,H,nQ,nK] or None; mask [B,1|H,nQ,nK]: bool
    entries set the *logit* to -100, float masks are added (:743-749).  Returns (o, probabilities after dropout), the
    tensor the reference returns as `attn` (:751-758)."""
    qh, kh, vh = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s.masked_fill(mask, -100.0) if mask.dtype == torch.bool else s + mask
    p = drop(torch.softmax(s, dim=-1))
    o = (p @ vh).permute(0, 2, 1, 3)
    return o, p


class MultiheadSelfAttention(nn.Module):
    """Stand-in for ``nn.MultiheadAttention(d_model, nhead, dropout)`` as used at :468/:541: same parameter
    names (in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias), sequence-first tensors,
    returns (output, None)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = Linear(embed_dim, embed_dim)
        self.dropout = dropout
        self.attn_drop = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, query, key, value, attn_mask=None, key_padding_mask=None, need_weights=False):
        L, B, D = query.shape
        H, hd = self.num_heads, self.head_dim
        wq, wk, wv = self.in_proj_weight.chunk(3)
        bq, bk, bv = self.in_proj_bias.chunk(3)
        if key is query:
            qk = ops.linear(query, self.in_proj_weight[: 2 * D], self.in_proj_bias[: 2 * D])
            q, k = qk[..., :D], qk[..., D:]
        else:
            q, k = ops.linear(query, wq, bq), ops.linear(key, wk, bk)
        v = ops.linear(value, wv, bv)
        q = (q * (hd ** -0.5)).reshape(L, B, H, hd).transpose(0, 1)
        k = k.reshape(-1, B, H, hd).transpose(0, 1)
        v = v.reshape(-1, B, H, hd).transpose(0, 1)
        plain = attn_mask is None and key_padding_mask is None and not need_weights
        weights = None
        if plain:
            o = ops.rpe_attention(q, k, v, dropout_p=self.dropout if self.training else 0.0)
        else:
            bias = None
            if attn_mask is not None:
                bias = torch.zeros_like(attn_mask, dtype=q.dtype).masked_fill(attn_mask, float("-inf")) \
                    if attn_mask.dtype == torch.bool else attn_mask
            if key_padding_mask is not None:
                pad = torch.zeros(B, 1, 1, k.shape[1], dtype=q.dtype, device=q.device).masked_fill(
                    key_padding_mask[:, None, None, :], float("-inf"))
                bias = pad if bias is None else bias + pad
            o, weights = _dense_attention(q, k, v, bias, self.attn_drop)
            weights = weights.mean(dim=1) if need_weights else None
        o = o.transpose(0, 1).reshape(L, B, D)
        return self.out_proj(o), weights


class ShareSelfAttention(nn.Module):
    """Self attention with one K/V head shared by all query heads (:609-653)."""

    def __init__(self, dim, num_heads, qkv_bias=True, dropout=0.0, args=None):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = Linear(dim, dim, bias=qkv_bias)
        self.k = Linear(dim, dim // num_heads, bias=qkv_bias)
        self.v = Linear(dim, dim // num_heads, bias=qkv_bias)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, query, key, value=None, attn_mask=None, key_padding_mask=None):
        assert attn_mask is None and key_padding_mask is None
        L, B, D = query.shape
        H = self.num_heads
        q = (self.q(query) * self.scale).reshape(L, B, H, D // H).transpose(0, 1)
        k = self.k(key).transpose(0, 1).unsqueeze(2)
        v = self.v(value).transpose(0, 1).unsqueeze(2)
        o = ops.rpe_attention(q, k, v, dropout_p=self.attn_drop.p if self.training else 0.0)
        x = self.proj(o.transpose(0, 1).reshape(L, B, D))
        return self.proj_drop(x), None


class GlobalShareCrossAttention(nn.Module):
    """Cross attention with the 3-D Vertex relative position bias (:656-758)."""

    def __init__(self, dim, num_heads, qkv_bias=True, attn_drop=0.0, proj_drop=0.0, args=None):
        super().__init__()
        assert dim % num_heads == 0, "dim should be divisible by num_heads"
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.log_scale = args.log_scale
        self.rpe_quant = args.rpe_quant
        self.angle_type = args.angle_type
        self.interp_method, max_value, num_points = self.rpe_quant.split("_")
        max_value, num_points = float(max_value), int(num_points)
        lin = torch.linspace(-max_value, max_value, num_points, dtype=torch.float32)
        self.register_buffer("relative_coords_table",
                             torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), dim=-1).unsqueeze(0))
        self.max_value = max_value
        self.cpb_mlps = get_clones(self.build_cpb_mlp(3, args.rpe_dim, num_heads), 8)
        self.q = Linear(dim, dim, bias=qkv_bias)
        self.k = Linear(dim, dim // num_heads, bias=qkv_bias)
        self.v = Linear(dim, dim // num_heads, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    def build_cpb_mlp(self, in_dim, hidden_dim, out_dim):
        return nn.Sequential(nn.Linear(in_dim, hidden_dim, bias=True), nn.ReLU(inplace=False),
                             nn.Linear(hidden_dim, out_dim, bias=False))

    def vertex_tables(self):
        """[8, n, n, n, H]: the eight per-vertex MLPs evaluated on the lattice (:725).  Kernel input; its
        gradient (dTables) is the kernel output that autograd carries on into the MLP weights.
        The eight MLPs (3 -> rpe_dim -> ReLU -> H) share their input, so they run as two batched GEMMs instead of sixteen
        tiny ones (and four instead of thirty-two in the backward): same parameters, same values."""
        mlps = list(self.cpb_mlps)
        lat = self.relative_coords_table
        n = lat.shape[1]
        if not all(isinstance(m, nn.Sequential) and len(m) == 3 and isinstance(m[0], nn.Linear) and isinstance(m[2], nn.Linear)
                   and m[0].bias is not None and m[2].bias is None for m in mlps):
            return torch.stack([mlp(lat)[0] for mlp in mlps])
        pts = lat.reshape(-1, lat.shape[-1])                                           # [n^3, 3]
        w1 = torch.stack([m[0].weight for m in mlps])                                  # [8, hid, 3]
        b1 = torch.stack([m[0].bias for m in mlps])                                    # [8, hid]
        w2 = torch.stack([m[2].weight for m in mlps])                                  # [8, H, hid]
        # (ops.batched_linear: bmm / baddbmm whose weight gradients come out in the parameters' layout)
        hid = torch.relu(ops.batched_linear(pts.unsqueeze(0).expand(len(mlps), -1, -1), w1, b1))
        return ops.batched_linear(hid, w2).view(len(mlps), n, n, n, -1)

    def forward(self, query, key, reference_point, reference_angle, xyz, attn_mask=None, key_padding_mask=None,
                need_weights=False):
        # query [nQ,B,D]  key [nK,B,D]  reference_point [B,nQ,8,3]  xyz [B,nK,3]
        # need_weights=True (GlobalDecoderLayer's return_attn_weights) or an attn_mask select the materialising path,
        # which returns the [B,H,nQ,nK] probabilities like the reference does; otherwise attn is None.
        nQ, B, D = query.shape
        H = self.num_heads
        if self.interp_method != "bilinear":
            raise NotImplementedError("only rpe_quant='bilinear_<max>_<n>' is implemented")
        rotate = self.angle_type == "object_coords" and reference_angle is not None
        tables = self.vertex_tables()
        q = (self.q(query) * self.scale).reshape(nQ, B, H, D // H).transpose(0, 1)
        k = self.k(key).transpose(0, 1).unsqueeze(2)
        v = self.v(key).transpose(0, 1).unsqueeze(2)
        ref = reference_point.detach().float()
        ang = reference_angle.detach().float() if rotate else None
        pts = xyz.detach().float()
        fused = attn_mask is None and not need_weights
        attn = None
        if fused:
            o = ops.rpe_attention(q, k, v, pts, ref, ang, tables, self.log_scale, self.max_value,
                                  dropout_p=self.attn_drop.p if self.training else 0.0)
        else:
            bias = _RpeBiasFn.apply(pts, ref, ang, tables, self.log_scale, self.max_value)
            o, attn = _dense_attention(q, k, v, bias, self.attn_drop, None if attn_mask is None else attn_mask.unsqueeze(1))
        x = self.proj(o.transpose(0, 1).reshape(nQ, B, D))
        return self.proj_drop(x), attn


class _RpeBiasFn(torch.autograd.Function):
    """Materialised bias with a table gradient (debug / mask / attention-dropout path only)."""

    @staticmethod
    def forward(ctx, xyz, ref, ang, tables, log_scale, max_value):
        ctx.save_for_backward(xyz, ref, ang, tables)
        ctx.meta = (log_scale, max_value)
        return ops.rpe_bias(xyz.contiguous(), ref.contiguous(), tables.contiguous(),
                            None if ang is None else ang.contiguous(), log_scale, max_value)

    @staticmethod
    def backward(ctx, g):
        xyz, ref, ang, tables = ctx.saved_tensors
        return None, None, None, ops.rpe_bias_grad_tables(xyz, ref, ang, tables, g.contiguous(), *ctx.meta), None, None


# ------------------------------------------------------------------------------------------------- layers
class FFNLayer(nn.Module):
    """Pre-norm feed forward block applied to the encoder tokens (:585-606)."""

    def __init__(self, d_model, dim_feedforward=256, dropout=0.1, norm_fn_name="ln", activation="relu",
                 normalize_before=True):
        super().__init__()
        self.linear1 = Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout, inplace=False)
        self.linear2 = Linear(dim_feedforward, d_model)
        self.norm = NORM_DICT[norm_fn_name](d_model)
        self.activation = ACTIVATION_DICT[activation]()
        self.normalize_before = normalize_before

    def forward_pre(self, memory):
        memory = self.norm(memory)
        return memory + self.dropout(self.linear2(self.dropout(self.activation(self.linear1(memory)))))

    def forward(self, memory):
        return self.forward_pre(memory)


class GlobalDecoderLayer(nn.Module):
    """Self attention over the queries, Vertex-RPE cross attention to the point tokens, FFN (:455-582)."""

    def __init__(self, d_model, nhead=4, dim_feedforward=256, dropout=0.1, dropout_attn=None, activation="relu",
                 normalize_before=True, norm_fn_name="ln", pos_for_key=False, args=None):
        super().__init__()
        if dropout_attn is None:
            dropout_attn = dropout
        self.pos_for_key = pos_for_key
        if args.share_selfattn:
            self.self_attn = ShareSelfAttention(d_model, nhead, dropout=dropout)
        else:
            self.self_attn = MultiheadSelfAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = GlobalShareCrossAttention(d_model, nhead, attn_drop=dropout, proj_drop=dropout, args=args)
        self.norm1 = NORM_DICT[norm_fn_name](d_model)
        self.norm2 = NORM_DICT[norm_fn_name](d_model)
        self.norm3 = NORM_DICT[norm_fn_name](d_model)
        self.dropout1 = nn.Dropout(dropout, inplace=False)
        self.dropout2 = nn.Dropout(dropout, inplace=False)
        self.dropout3 = nn.Dropout(dropout, inplace=False)
        self.linear1 = Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout, inplace=False)
        self.linear2 = Linear(dim_feedforward, d_model)
        self.activation = ACTIVATION_DICT[activation]()
        self.normalize_before = normalize_before

    @staticmethod
    def with_pos_embed(tensor, pos: Optional[Tensor]):
        return tensor if pos is None else tensor + pos

    def _cross(self, tgt, memory, reference_point, reference_angle, enc_xyz, memory_mask, memory_key_padding_mask, pos,
               query_pos, need_weights=False):
        key = self.with_pos_embed(memory, pos) if self.pos_for_key else memory
        return self.multihead_attn(query=self.with_pos_embed(tgt, query_pos), key=key, reference_point=reference_point,
                                   reference_angle=reference_angle, xyz=enc_xyz, attn_mask=memory_mask,
                                   key_padding_mask=memory_key_padding_mask, need_weights=need_weights)

    def forward_post(self, tgt, memory, reference_point, reference_angle, enc_xyz, point_cloud_dims, tgt_mask=None,
                     memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None,
                     return_attn_weights=False):
        qk = self.with_pos_embed(tgt, query_pos)
        tgt = self.norm1(tgt + self.dropout1(self.self_attn(qk, qk, value=tgt, attn_mask=tgt_mask,
                                                            key_padding_mask=tgt_key_padding_mask)[0]))
        x, attn = self._cross(tgt, memory, reference_point, reference_angle, enc_xyz, memory_mask,
                              memory_key_padding_mask, pos, query_pos, return_attn_weights)
        tgt = self.norm2(tgt + self.dropout2(x))
        tgt = self.norm3(tgt + self.dropout3(self.linear2(self.dropout(self.activation(self.linear1(tgt))))))
        return (tgt, attn) if return_attn_weights else (tgt, None)

    def forward_pre(self, tgt, memory, reference_point, reference_angle, enc_xyz, point_cloud_dims, tgt_mask=None,
                    memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None,
                    return_attn_weights=False):
        t2 = self.norm1(tgt)
        qk = self.with_pos_embed(t2, query_pos)
        tgt = tgt + self.dropout1(self.self_attn(qk, qk, value=t2, attn_mask=tgt_mask,
                                                 key_padding_mask=tgt_key_padding_mask)[0])
        t2 = self.norm2(tgt)
        x, attn = self._cross(t2, memory, reference_point, reference_angle, enc_xyz, memory_mask,
                              memory_key_padding_mask, pos, query_pos, return_attn_weights)
        tgt = tgt + self.dropout2(x)
        t2 = self.norm3(tgt)
        tgt = tgt + self.dropout3(self.linear2(self.dropout(self.activation(self.linear1(t2)))))
        return (tgt, attn) if return_attn_weights else (tgt, None)

    def forward(self, tgt, memory, reference_point, reference_angle, enc_xyz, point_cloud_dims, tgt_mask=None,
                memory_mask=None, tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None,
                return_attn_weights=False):
        fn = self.forward_pre if self.normalize_before else self.forward_post
        return fn(tgt, memory, reference_point, reference_angle, enc_xyz, point_cloud_dims, tgt_mask, memory_mask,
                  tgt_key_padding_mask, memory_key_padding_mask, pos, query_pos, return_attn_weights)


class TransformerDecoder(nn.Module):
    """Proposal stage on the encoder tokens + ``num_layers`` refinement layers with box heads (:105-452)."""

    def __init__(self, first_layer, decoder_layer, dataset_config, num_layers, decoder_dim=256, mlp_dropout=0.3,
                 mlp_norm="bn1d", mlp_act="relu", mlp_sep=False, pos_for_key=False, num_queries=256, cls_loss="celoss",
                 norm_fn_name="ln", is_bilable=False, q_content="sample", return_intermediate=False,
                 weight_init_name="xavier_uniform", args=None):
        super().__init__()
        self.first_layer = first_layer
        self.layers = get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.dec_output_dim = self.layers[0].linear2.out_features
        self.norm = NORM_DICT[norm_fn_name](self.dec_output_dim) if norm_fn_name is not None else None
        self.is_bilable = is_bilable
        self.pos_for_key = pos_for_key
        self.num_queries = num_queries
        self.q_content = q_content
        self.query_pos_projection = nn.ModuleList(
            [PositionEmbeddingLearned(6, self.dec_output_dim) for _ in range(num_layers)])
        if pos_for_key:
            self.key_pos_projection = nn.ModuleList(
                [PositionEmbeddingLearned(3, self.dec_output_dim) for _ in range(num_layers)])
        if q_content in ("random", "random_add"):
            self.query_embed = nn.Embedding(num_queries, self.dec_output_dim)
        self.return_intermediate = return_intermediate
        self._reset_parameters(weight_init_name)
        self.mlp_norm, self.mlp_act, self.mlp_sep, self.cls_loss = mlp_norm, mlp_act, mlp_sep, cls_loss
        self.build_mlp_heads(dataset_config, decoder_dim, mlp_dropout)
        self.build_pointcls_heads(dataset_config, decoder_dim, mlp_dropout)
        head_sets = list(self.mlp_heads) if mlp_sep else [self.mlp_heads]
        if cls_loss.split("_")[0] == "focalloss":
            prior = -math.log((1 - 0.01) / 0.01)
            for hs in head_sets:
                last = hs["sem_cls_head"].layers[-1]
                last.bias.data = torch.ones(last.bias.shape[0]) * prior
        for hs in head_sets:
            for name in ("center_head", "size_head"):
                nn.init.constant_(hs[name].layers[-1].weight.data, 0.0)
                nn.init.constant_(hs[name].layers[-1].bias.data, 0.0)
        self.box_processor = BoxProcessor(dataset_config, cls_loss=cls_loss)
        self.sort_keys = True       # Morton-order the key tokens once per forward (see module docstring)
        self.parallel_heads = os.environ.get("VDETR_B200_PARALLEL_HEADS", "1") != "0"
        self.group_heads = os.environ.get("VDETR_B200_GROUP_HEADS", "1") != "0"
        self.batch_first_memory = os.environ.get("VDETR_B200_BATCH_FIRST", "1") != "0"   # see ops.batch_first_backed
        self.fuse_box_decode = os.environ.get("VDETR_B200_FUSE_BOX_DECODE", "1") != "0"
        # the fused box decode hard-codes utils/box_util.py:294-358 (what both dataset configs of the reference use)
        self.standard_corners = type(dataset_config).__name__ in ("ScanNetBoxConfig", "ScannetDatasetConfig", "SunrgbdDatasetConfig")

    def _head_factory(self, decoder_dim, mlp_dropout):
        return partial(GenericMLP, norm_fn_name=self.mlp_norm, activation=self.mlp_act, use_conv=True,
                       hidden_dims=[decoder_dim, decoder_dim], dropout=mlp_dropout, input_dim=decoder_dim)

    def build_pointcls_heads(self, dataset_config, decoder_dim, mlp_dropout):
        extra = 0 if self.cls_loss.split("_")[0] == "focalloss" else 1
        self.pointcls_heads = self._head_factory(decoder_dim, mlp_dropout)(output_dim=dataset_config.num_semcls + extra)

    def build_mlp_heads(self, dataset_config, decoder_dim, mlp_dropout):
        mk = self._head_factory(decoder_dim, mlp_dropout)
        extra = 0 if self.cls_loss.split("_")[0] == "focalloss" else 1
        heads = [("sem_cls_head", mk(output_dim=dataset_config.num_semcls + extra)),
                 ("center_head", mk(output_dim=3)), ("size_head", mk(output_dim=3)),
                 ("angle_cls_head", mk(output_dim=dataset_config.num_angle_bin)),
                 ("angle_residual_head", mk(output_dim=dataset_config.num_angle_bin))]
        if not self.mlp_sep:
            self.mlp_heads = nn.ModuleDict(heads)
        elif self.is_bilable:
            self.mlp_heads = get_clones(nn.ModuleDict(heads), self.num_layers)
            first = copy.deepcopy(heads)
            first[0] = ("sem_cls_head", mk(output_dim=1))
            self.mlp_heads.insert(0, nn.ModuleDict(first))
        else:
            self.mlp_heads = get_clones(nn.ModuleDict(heads), self.num_layers + 1)

    def _reset_parameters(self, weight_init_name):
        init = WEIGHT_INIT_DICT[weight_init_name]
        for _, p in self.named_parameters():
            if p.dim() > 1:
                init(p)

    HEAD_NAMES = ("sem_cls_head", "center_head", "size_head", "angle_cls_head", "angle_residual_head")

    def _grouped_plan(self, heads):
        """The 5 heads of a level as one stack when they share the reference's default structure
        [Conv1d(k=1) - BatchNorm1d - ReLU - Dropout] x 2 - Conv1d(k=1) (models/helpers.py:74-141 with use_conv, bn1d, relu);
        None otherwise."""
        stacks = [list(heads[n].layers) for n in self.HEAD_NAMES]
        kinds = (nn.Conv1d, nn.BatchNorm1d, nn.ReLU, nn.Dropout, nn.Conv1d, nn.BatchNorm1d, nn.ReLU, nn.Dropout, nn.Conv1d)
        for st in stacks:
            if len(st) != len(kinds) or any(type(m) is not k for m, k in zip(st, kinds)):
                return None
            if any(m.kernel_size != (1,) or m.stride != (1,) or m.groups != 1 for m in (st[0], st[4], st[8])):
                return None
        c0 = stacks[0]
        for st in stacks:
            if (st[0].weight.shape != c0[0].weight.shape or st[4].weight.shape != c0[4].weight.shape
                    or (st[0].bias is None) != (c0[0].bias is None) or (st[4].bias is None) != (c0[4].bias is None)
                    or st[8].bias is None or st[3].p != c0[3].p or st[7].p != c0[7].p):
                return None
        return stacks

    @staticmethod
    def _bn_relu_group(x, bns, layout, drop):
        """drop(relu(bn_g(x_g))) for the heads' BatchNorm layers and the nn.Dropout behind them: the library's grouped kernels
        in training (dropout fused), the running statistics as a per-channel affine map otherwise."""
        if ops.bn_relu_train_group_supported(x, bns, layout):
            return ops.bn_relu_train_group(x, bns, layout, drop.p if drop.training else 0.0)
        return drop(TransformerDecoder._bn_relu_group_plain(x, bns, layout))

    @staticmethod
    def _bn_relu_group_plain(x, bns, layout):
        if any(b.training for b in bns):
            # widths the kernels do not cover: per-head stock BatchNorm
            C = bns[0].num_features
            parts = x.split(C, dim=1) if layout == "cl" else x.unbind(0)
            ys = [F.relu(b(p)) for b, p in zip(bns, parts)]
            return torch.cat(ys, dim=1) if layout == "cl" else torch.stack(ys)
        scale = torch.stack([b.weight * torch.rsqrt(b.running_var + b.eps) for b in bns])           # [G, C]
        shift = torch.stack([b.bias for b in bns]) - torch.stack([b.running_mean for b in bns]) * scale
        if layout == "cl":
            return F.relu(x * scale.reshape(1, -1) + shift.reshape(1, -1))
        return F.relu(x * scale.unsqueeze(1) + shift.unsqueeze(1))

    def _run_heads_grouped(self, stacks, feats, nQ, B, batch_first=False):
        """feats [T = nQ*B, C] -> {head: [B,nQ,out]}: layer 1 of all heads is ONE GEMM against the concatenated weights
        ([T, G*C], channels last), layers 2 and 3 are batched GEMMs over the heads ([G, T, C]); each BatchNorm+ReLU level
        is one grouped kernel pair.  ~4x fewer launches than head-by-head evaluation, same parameters and results."""
        G, T = len(stacks), feats.shape[0]
        C = stacks[0][0].out_channels
        W1 = torch.cat([st[0].weight.squeeze(-1) for st in stacks])                                 # [G*C, Cin]
        b1 = torch.cat([st[0].bias for st in stacks]) if stacks[0][0].bias is not None else None
        h = ops.linear(feats, W1, b1)                                                               # [T, G*C]
        h = self._bn_relu_group(h, [st[1] for st in stacks], "cl", stacks[0][3])
        W2 = torch.stack([st[4].weight.squeeze(-1) for st in stacks])                               # [G, C, C]
        b2 = torch.stack([st[4].bias for st in stacks]) if stacks[0][4].bias is not None else None
        h = ops.batched_linear(h.view(T, G, C).transpose(0, 1), W2, b2)                             # [G, T, C]
        h = self._bn_relu_group(h, [st[5] for st in stacks], "gm", stacks[0][7])
        outs = [st[8].out_channels for st in stacks]
        omax = (max(outs) + 7) // 8 * 8        # multiple of 8: cuBLAS otherwise falls back to its unaligned legacy kernels (140 us per call)
        W3 = _PadStack.apply(omax, *[st[8].weight.squeeze(-1) for st in stacks])                    # [G, omax, C]
        b3 = _PadStack.apply(omax, *[st[8].bias for st in stacks])                                  # [G, omax]
        o = ops.batched_linear(h, W3, b3)                                                           # [G, T, omax]
        return dict(zip(self.HEAD_NAMES, _SplitHeadOutputs.apply(o, nQ, B, outs, batch_first)))

    def _run_heads(self, heads, box_features):
        """box_features [nQ,B,C] -> {head: [B,nQ,out]} (:256-300).  The heads are evaluated token-major (GEMMs on the
        contiguous [nQ*B, C] rows instead of Conv1d on a permuted copy, see helpers.pointwise_tokens).  Heads of the
        reference's default structure run as one grouped stack (_run_heads_grouped); otherwise the 5 heads of a level,
        which read the same features and are independent, are issued on 5 forked streams and joined."""
        nQ, B, C = box_features.shape
        tokens = all(heads[n].supports_tokens for n in self.HEAD_NAMES)
        bf = tokens and ops.batch_first_backed(box_features)       # tokens in memory order (b, q): no transposing copy
        if tokens and self.group_heads and box_features.is_cuda:
            stacks = self._grouped_plan(heads)
            if stacks is not None:
                feats = box_features.transpose(0, 1).reshape(B * nQ, C) if bf else box_features.reshape(nQ * B, C)
                return self._run_heads_grouped(stacks, feats, nQ, B, bf)
        if bf:
            feats = box_features.transpose(0, 1).reshape(B * nQ, C)
        else:
            feats = box_features.reshape(nQ * B, C) if tokens else box_features.permute(1, 2, 0)

        def run(n):
            if bf:
                return heads[n].forward_tokens(feats).view(B, nQ, -1)
            if tokens:
                return heads[n].forward_tokens(feats).view(nQ, B, -1).transpose(0, 1)
            return heads[n](feats).transpose(1, 2)
        if not (feats.is_cuda and self.parallel_heads) or ops.SYNC_BN_ACTIVE:
            # (SyncBatchNorm: the exchanges inside the BatchNorm launches share one flag channel -- one stream, one order)
            return {n: run(n) for n in self.HEAD_NAMES}
        cur = torch.cuda.current_stream(feats.device)
        streams = _head_streams(feats.device, len(self.HEAD_NAMES))
        out = {}
        for n, st in zip(self.HEAD_NAMES, streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                out[n] = run(n)
        for n, st in zip(self.HEAD_NAMES, streams):
            cur.wait_stream(st)
            out[n].record_stream(cur)        # allocated on a side stream, consumed (and later freed) on `cur`
        return out

    def get_proposal_box_predictions_refine(self, idx, query_xyz, point_cloud_dims, box_features,
                                            pre_center_normalized=None, pre_size_normalized=None):
        """box_features [nQ,B,C] -> dict of box predictions (:244-333)."""
        assert pre_center_normalized is not None and pre_size_normalized is not None
        heads = self.mlp_heads[idx] if self.mlp_sep else self.mlp_heads
        lo, hi = point_cloud_dims
        scene = (hi - lo).unsqueeze(1)
        origin = lo.unsqueeze(1)
        raw = self._run_heads(heads, box_features)
        cls_logits = raw["sem_cls_head"]
        center_reg = raw["center_head"].contiguous()
        size_reg = raw["size_head"].contiguous()
        angle_logits = raw["angle_cls_head"]
        angle_res_norm = raw["angle_residual_head"]
        angle_res = angle_res_norm * (np.pi / angle_res_norm.shape[-1])
        angle, angle_prob = self.box_processor.compute_predicted_angle(angle_logits, angle_res)
        ref_lidar = None
        if (self.fuse_box_decode and angle_logits.shape[-1] == 1 and self.standard_corners
                and ops.box_decode_supported(center_reg, size_reg, pre_center_normalized, pre_size_normalized, lo, hi)):
            # num_angle_bin == 1: the angle is identically zero (:49-59), rotated and axis-aligned corners coincide, and the
            # whole decode is one kernel (csrc/boxdecode.cu); it also leaves the next layer's lidar-frame reference_point
            center, center_norm, size, size_norm, corners, pre_center, pre_size, ref_lidar = ops.box_decode(
                center_reg, size_reg, pre_center_normalized, pre_size_normalized, lo, hi)
            corners0 = corners
        else:
            pre_center = pre_center_normalized * scene + origin
            pre_size = pre_size_normalized * scene
            center = center_reg * pre_size + pre_center
            center_norm = (center - origin) / scene
            size = torch.exp(size_reg) * pre_size
            size_norm = size / scene
            corners = self.box_processor.box_parametrization_to_corners(center, size, angle)
            angle0, _ = self.box_processor.compute_predicted_angle(angle_logits, angle_res, zero_angle=True)
            corners0 = self.box_processor.box_parametrization_to_corners(center, size, angle0)
        with torch.no_grad():
            semcls_prob, objectness = self.box_processor.compute_objectness_and_cls_prob(cls_logits)
        extra = {} if ref_lidar is None else {"_reference_point_lidar": ref_lidar}      # popped by forward()
        return {**extra, "sem_cls_logits": cls_logits, "center_normalized": center_norm.contiguous(),
                "center_unnormalized": center, "size_normalized": size_norm, "size_unnormalized": size,
                "angle_logits": angle_logits, "angle_prob": angle_prob, "angle_residual": angle_res,
                "angle_residual_normalized": angle_res_norm, "angle_continuous": angle,
                "objectness_prob": objectness, "sem_cls_prob": semcls_prob, "box_corners": corners,
                "box_corners_axis_align": corners0, "pre_box_center_unnormalized": pre_center,
                "center_reg": center_reg, "pre_box_size_unnormalized": pre_size, "size_reg": size_reg}

    def forward(self, tgt, memory, query_xyz, enc_xyz, point_cloud_dims, tgt_mask=None, memory_mask=None,
                tgt_key_padding_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None, transpose_swap=False,
                return_attn_weights=False, enc_box_predictions=None, enc_box_features=None):
        preds, attns = [], []
        output = self.first_layer(enc_box_features)
        pred = self.get_proposal_box_predictions_refine(
            0, query_xyz, point_cloud_dims, self.norm(output),
            pre_center_normalized=enc_box_predictions["center_normalized"],
            pre_size_normalized=enc_box_predictions["size_normalized"])
        pred.pop("_reference_point_lidar", None)
        ref_lidar = None
        if self.return_intermediate:
            preds.append(pred)
        score = pred["objectness_prob"].clone().detach()
        B, nprop = score.shape
        if nprop >= self.num_queries:
            top = torch.topk(score, self.num_queries, dim=1)[1]
        else:
            top = torch.arange(nprop, device=score.device).unsqueeze(0).repeat(B, 1)

        def pick(t):
            idx = top.reshape(top.shape + (1,) * (t.dim() - 2)).expand(top.shape + t.shape[2:])
            return torch.gather(t.clone().detach(), 1, idx)
        reference_point = convert_corners_camera2lidar(pick(pred["box_corners"]))
        reference_center = pick(pred["center_unnormalized"])
        query_xyz = reference_center.clone().detach()
        reference_size = pick(pred["size_unnormalized"])
        reference_angle = pick(pred["angle_continuous"])
        proposal_center_normalized = pick(pred["center_normalized"])
        proposal_size_normalized = pick(pred["size_normalized"])
        output = torch.gather(output.permute(1, 0, 2), 1, top.unsqueeze(-1).expand(-1, -1, output.shape[-1])) \
            .permute(1, 0, 2).contiguous()
        if self.q_content == "zero":
            output = torch.zeros_like(output)
        elif self.q_content == "random":
            output = self.query_embed.weight.unsqueeze(1).repeat(1, output.shape[1], 1)
        elif self.q_content == "random_add":
            output = output + self.query_embed.weight.unsqueeze(1).repeat(1, output.shape[1], 1)

        # keys in Morton order for the cross-attention kernels (results are permutation invariant over keys)
        mem_l, xyz_l, perm = memory, enc_xyz, None
        if self.sort_keys and memory_mask is None and not return_attn_weights and not self.pos_for_key:
            perm = morton_order(enc_xyz.detach())
            xyz_l = torch.gather(enc_xyz, 1, perm.unsqueeze(-1).expand(-1, -1, 3))
            # gathered straight into batch-first memory and handed on as a [nK, B, C] view of it (ops.batch_first_backed)
            mem_l = torch.gather(memory.transpose(0, 1), 1, perm.unsqueeze(-1).expand(-1, -1, memory.shape[-1])).transpose(0, 1)
        if self.batch_first_memory:
            # token activations live in batch-first memory behind the reference's [nQ, B, C] shapes: the attention kernels
            # want [B, nQ, H, hd], and a sequence-first memory order costs them 4 transposing copies per call and direction
            output = output.transpose(0, 1).contiguous().transpose(0, 1)

        for idx, layer in enumerate(self.layers):
            if idx > 0:
                reference_point = ref_lidar if ref_lidar is not None else \
                    convert_corners_camera2lidar(pred["box_corners"].clone().detach())
                reference_center = pred["center_unnormalized"].clone().detach()
                reference_size = pred["size_unnormalized"].clone().detach()
                reference_angle = pred["angle_continuous"].clone().detach()
            query_reference = torch.cat([reference_center, reference_size], dim=-1)
            qpos = self.query_pos_projection[idx].forward_tokens(query_reference)
            if self.pos_for_key:
                pos = self.key_pos_projection[idx](enc_xyz).permute(2, 0, 1)
            output, attn = layer(output, mem_l, reference_point, reference_angle, xyz_l, point_cloud_dims,
                                 tgt_mask=tgt_mask, memory_mask=memory_mask, tgt_key_padding_mask=tgt_key_padding_mask,
                                 memory_key_padding_mask=memory_key_padding_mask, pos=pos, query_pos=qpos,
                                 return_attn_weights=return_attn_weights)
            pred = self.get_proposal_box_predictions_refine(
                idx + 1, query_xyz, point_cloud_dims, self.norm(output),
                pre_center_normalized=proposal_center_normalized, pre_size_normalized=proposal_size_normalized)
            ref_lidar = pred.pop("_reference_point_lidar", None)      # lidar-frame corners left by the fused box decode
            if self.return_intermediate:
                preds.append(pred)
            if return_attn_weights:
                attns.append(attn)
        if return_attn_weights:
            attns = torch.stack(attns)
        if self.return_intermediate:
            return {"outputs": preds[-1], "aux_outputs": preds[:-1]}, attns
        return {"outputs": pred}, attns
