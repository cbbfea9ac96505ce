"""CPU oracle: plain-PyTorch fp32 restatement of the V-DETR decoder hot path.

TEST INFRASTRUCTURE ONLY (see oracle/rpe_attention.py header).  This file is also the ``port`` arm
of ``bench.py --impl reference`` / ``cpu_baseline`` (the reference's own code is Python and cannot
travel to the GPU box, so its algorithm is restated here and pinned with golden vectors).

Parity status: PINNED by tests/golden/*.npz, generated from the unmodified reference by
tests/golden/make_golden.py (``tests/test_oracle_golden.py``).

Module/parameter names follow the reference state_dict so one set of weights
loads into the reference, this oracle and the CUDA product alike.

Reference files restated (all under /root/reference):
  models/vdetr_transformer.py:20-90    BoxProcessor           -> _angle_from_heads, _objectness
  models/vdetr_transformer.py:105-452  TransformerDecoder     -> OracleDecoder
  models/vdetr_transformer.py:455-582  GlobalDecoderLayer     -> OracleDecoderLayer (pre/post norm)
  models/vdetr_transformer.py:585-606  FFNLayer               -> OracleFFN
  models/vdetr_transformer.py:609-653  ShareSelfAttention     -> OracleSharedSelfAttention
  models/vdetr_transformer.py:656-758  GlobalShareCrossAttention -> OracleVertexRPECrossAttention
  models/helpers.py:17-33,74-141       PositionEmbeddingLearned, GenericMLP
  utils/box_util.py:294-358, datasets/scannet.py:168-171   box corners
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- box geometry
class OracleBoxConfig:
    """The three things the decoder reads from ScannetDatasetConfig (datasets/scannet.py:38-41,168-171)."""

    def __init__(self, num_semcls: int = 18, num_angle_bin: int = 1):
        self.num_semcls = num_semcls
        self.num_angle_bin = num_angle_bin

    @staticmethod
    def box_parametrization_to_corners(center, size, angle):
        return corners_camera_frame(center, size, angle)


def corners_camera_frame(center, size, angle):
    """utils/box_util.py:319-358 applied to flip_axis_to_camera_tensor(center) (:294-301)."""
    cam_center = torch.stack([center[..., 0], -center[..., 2], center[..., 1]], dim=-1)
    half_l, half_w, half_h = size[..., 0:1] / 2, size[..., 1:2] / 2, size[..., 2:3] / 2
    sx = center.new_tensor([1, 1, -1, -1, 1, 1, -1, -1])
    sy = center.new_tensor([1, 1, 1, 1, -1, -1, -1, -1])
    sz = center.new_tensor([1, -1, -1, 1, 1, -1, -1, 1])
    local = torch.stack([half_l * sx, half_h * sy, half_w * sz], dim=-1)      # [...,8,3]
    c, s = torch.cos(angle), torch.sin(angle)
    rot = torch.zeros(angle.shape + (3, 3), dtype=torch.float32, device=angle.device)
    rot[..., 0, 0] = c
    rot[..., 0, 2] = s
    rot[..., 1, 1] = 1
    rot[..., 2, 0] = -s
    rot[..., 2, 2] = c
    return torch.matmul(local, rot.transpose(-1, -2)) + cam_center.unsqueeze(-2)


def camera_to_world_corners(corners_cam):
    """models/vdetr_transformer.py:98-102: (x_c, y_c, z_c) -> (x_c, z_c, -y_c)."""
    return torch.stack([corners_cam[..., 0], corners_cam[..., 2], -corners_cam[..., 1]], dim=-1)


# ----------------------------------------------------------------------------- small modules
def conv_mlp(in_dim, hidden, out_dim, dropout):
    """GenericMLP(use_conv=True, norm='bn1d', act='relu', hidden_use_bias=False) (models/helpers.py:74-141)."""
    layers, prev = [], in_dim
    for h in hidden:
        layers += [nn.Conv1d(prev, h, 1, bias=False), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(dropout)]
        prev = h
    layers.append(nn.Conv1d(prev, out_dim, 1, bias=True))
    return nn.Sequential(*layers)


class _Head(nn.Module):
    def __init__(self, in_dim, out_dim, dropout):
        super().__init__()
        self.layers = conv_mlp(in_dim, [in_dim, in_dim], out_dim, dropout)

    def forward(self, x):
        return self.layers(x)


class OracleQueryPos(nn.Module):
    """PositionEmbeddingLearned (models/helpers.py:17-33)."""

    def __init__(self, in_ch, feats):
        super().__init__()
        self.position_embedding_head = nn.Sequential(
            nn.Conv1d(in_ch, feats, 1), nn.BatchNorm1d(feats), nn.ReLU(inplace=True),
            nn.Conv1d(feats, feats, 1))

    def forward(self, x):
        return self.position_embedding_head(x.transpose(1, 2).contiguous())


class OracleFFN(nn.Module):
    def __init__(self, d_model, dim_ff=256, dropout=0.1):
        super().__init__()
        self.linear1 = nn.Linear(d_model, dim_ff)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_ff, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x):
        y = self.norm(x)
        return y + self.dropout(self.linear2(self.dropout(F.relu(self.linear1(y)))))


class OracleSelfAttention(nn.Module):
    """nn.MultiheadAttention(d_model, nhead) as called at models/vdetr_transformer.py:468,541
    (same parameter names: in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, dim, heads, dropout=0.0):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = nn.Linear(dim, dim)
        self.drop = nn.Dropout(dropout)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def forward(self, query, key, value):
        L, B, D = query.shape
        H, hd = self.heads, D // self.heads
        wq, wk, wv = self.in_proj_weight.chunk(3)
        bq, bk, bv = self.in_proj_bias.chunk(3)
        q = F.linear(query, wq, bq).view(L, B, H, hd).permute(1, 2, 0, 3) * (hd ** -0.5)
        k = F.linear(key, wk, bk).view(-1, B, H, hd).permute(1, 2, 0, 3)
        v = F.linear(value, wv, bv).view(-1, B, H, hd).permute(1, 2, 0, 3)
        p = self.drop(torch.softmax(q @ k.transpose(-1, -2), dim=-1))
        o = (p @ v).permute(2, 0, 1, 3).reshape(L, B, D)
        return self.out_proj(o)


class OracleSharedSelfAttention(nn.Module):
    """ShareSelfAttention (models/vdetr_transformer.py:609-653): one K/V head shared by all heads."""

    def __init__(self, dim, heads, qkv_bias=True, dropout=0.0):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim // heads, bias=qkv_bias)
        self.v = nn.Linear(dim, dim // heads, bias=qkv_bias)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout)

    def forward(self, query, key, value):
        L, B, D = query.shape
        H = self.num_heads
        q = self.q(query).view(L, B, H, D // H).permute(1, 2, 0, 3) * self.scale
        k = self.k(key).permute(1, 0, 2).unsqueeze(1)
        v = self.v(value).permute(1, 0, 2).unsqueeze(1)
        p = self.attn_drop(torch.softmax(q @ k.transpose(-1, -2), dim=-1))
        o = (p @ v).permute(2, 0, 1, 3).reshape(L, B, D)
        return self.proj_drop(self.proj(o))


def rpe_bias_torch(ref_pts, xyz, tables, ref_angle=None, log_scale=512.0, max_value=4.0):
    """models/vdetr_transformer.py:708-731 with an explicit 8-corner gather (no grid_sample).
    ref_pts [B,nQ,8,3], xyz [B,nK,3], tables [8,N,N,N,H] -> [B,H,nQ,nK]."""
    N = tables.shape[1]
    out = 0
    for i in range(8):
        d = ref_pts[:, :, None, i, :] - xyz[:, None, :, :]
        if ref_angle is not None:
            c = torch.cos(ref_angle)[:, :, None]
            s = torch.sin(ref_angle)[:, :, None]
            d = torch.stack([c * d[..., 0] - s * d[..., 1], s * d[..., 0] + c * d[..., 1], d[..., 2]], -1)
        g = torch.sign(d) * torch.log2(d.abs() * log_scale + 1.0) / 3.0 / max_value
        p = ((g + 1) * N - 1) / 2
        p0 = torch.floor(p)
        f = p - p0
        i0 = p0.long()
        T = tables[i]
        acc = 0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    xc, yb, za = i0[..., 0] + dx, i0[..., 1] + dy, i0[..., 2] + dz
                    w = (f[..., 0] if dx else 1 - f[..., 0]) * (f[..., 1] if dy else 1 - f[..., 1]) \
                        * (f[..., 2] if dz else 1 - f[..., 2])
                    ok = (xc >= 0) & (xc < N) & (yb >= 0) & (yb < N) & (za >= 0) & (za < N)
                    w = torch.where(ok, w, torch.zeros_like(w))
                    acc = acc + w[..., None] * T[za.clamp(0, N - 1), yb.clamp(0, N - 1), xc.clamp(0, N - 1)]
        out = out + acc
    return out.permute(0, 3, 1, 2)


def rpe_bias_grid_sample(ref_pts, xyz, tables, ref_angle=None, log_scale=512.0, max_value=4.0):
    """Same quantity through F.grid_sample, the primitive the reference itself calls
    (models/vdetr_transformer.py:725-731).  Used for the CPU baseline timing (it is what the reference's CPU
    path executes); tests/test_oracle_golden.py checks that it agrees with the explicit gather above."""
    B, nQ = ref_pts.shape[:2]
    nK = xyz.shape[1]
    out = None
    for i in range(8):
        d = ref_pts[:, :, None, i, :] - xyz[:, None, :, :]
        if ref_angle is not None:
            c = torch.cos(ref_angle)[:, :, None]
            s = torch.sin(ref_angle)[:, :, None]
            d = torch.stack([c * d[..., 0] - s * d[..., 1], s * d[..., 0] + c * d[..., 1], d[..., 2]], -1)
        g = torch.sign(d) * torch.log2(d.abs() * log_scale + 1.0) / 3.0 / max_value
        tab = tables[i].permute(3, 0, 1, 2).unsqueeze(0)                              # [1,H,n,n,n]
        r = F.grid_sample(tab, g.reshape(1, 1, 1, -1, 3), mode="bilinear", padding_mode="zeros", align_corners=False)
        r = r.reshape(-1, B, nQ, nK).permute(1, 0, 2, 3)
        out = r if out is None else out + r
    return out


class OracleVertexRPECrossAttention(nn.Module):
    """GlobalShareCrossAttention (models/vdetr_transformer.py:656-758)."""

    def __init__(self, dim, heads, attn_drop=0.0, proj_drop=0.0, log_scale=512.0,
                 rpe_quant="bilinear_4_10", rpe_dim=128, angle_type=""):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.log_scale = log_scale
        self.angle_type = angle_type
        _, mv, npts = rpe_quant.split("_")
        self.max_value, self.num_points = float(mv), int(npts)
        lin = torch.linspace(-self.max_value, self.max_value, self.num_points)
        grid = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), dim=-1).unsqueeze(0)
        self.register_buffer("relative_coords_table", grid)
        one = nn.Sequential(nn.Linear(3, rpe_dim), nn.ReLU(), nn.Linear(rpe_dim, heads, bias=False))
        self.cpb_mlps = nn.ModuleList([copy.deepcopy(one) for _ in range(8)])
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim // heads)
        self.v = nn.Linear(dim, dim // heads)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.use_grid_sample = False

    def tables(self):
        return torch.stack([m(self.relative_coords_table)[0] for m in self.cpb_mlps])     # [8,N,N,N,H]

    def forward(self, query, key, reference_point, reference_angle, xyz, attn_mask=None):
        nQ, B, D = query.shape
        H = self.num_heads
        ang = reference_angle if (self.angle_type == "object_coords" and reference_angle is not None) else None
        fn = rpe_bias_grid_sample if self.use_grid_sample else rpe_bias_torch
        bias = fn(reference_point, xyz, self.tables(), ang, self.log_scale, self.max_value)
        q = self.q(query).view(nQ, B, H, D // H).permute(1, 2, 0, 3) * self.scale
        k = self.k(key).permute(1, 0, 2).unsqueeze(1)
        v = self.v(key).permute(1, 0, 2).unsqueeze(1)
        s = q @ k.transpose(-1, -2) + bias
        if attn_mask is not None:
            m = attn_mask.unsqueeze(1)
            s = s.masked_fill(m, -100.0) if m.dtype == torch.bool else s + m
        p = torch.softmax(s, dim=-1)
        o = (self.attn_drop(p) @ v).permute(2, 0, 1, 3).reshape(nQ, B, D)
        return self.proj_drop(self.proj(o)), p


class OracleDecoderLayer(nn.Module):
    def __init__(self, d_model=256, nhead=4, dim_ff=256, dropout=0.1, normalize_before=True,
                 share_selfattn=False, **rpe_kw):
        super().__init__()
        self.self_attn = (OracleSharedSelfAttention(d_model, nhead, dropout=dropout) if share_selfattn
                          else OracleSelfAttention(d_model, nhead, dropout=dropout))
        self.multihead_attn = OracleVertexRPECrossAttention(d_model, nhead, attn_drop=dropout,
                                                            proj_drop=dropout, **rpe_kw)
        self.norm1, self.norm2, self.norm3 = (nn.LayerNorm(d_model) for _ in range(3))
        self.dropout1, self.dropout2, self.dropout3, self.dropout = (nn.Dropout(dropout) for _ in range(4))
        self.linear1 = nn.Linear(d_model, dim_ff)
        self.linear2 = nn.Linear(dim_ff, d_model)
        self.normalize_before = normalize_before

    def forward(self, tgt, memory, reference_point, reference_angle, enc_xyz, query_pos, pos=None):
        def addpos(t, p):
            return t if p is None else t + p
        if self.normalize_before:                              # forward_pre, :531-568
            t2 = self.norm1(tgt)
            qk = addpos(t2, query_pos)
            tgt = tgt + self.dropout1(self.self_attn(qk, qk, t2))
            t2 = self.norm2(tgt)
            x, attn = self.multihead_attn(addpos(t2, query_pos), addpos(memory, pos), reference_point,
                                          reference_angle, enc_xyz)
            tgt = tgt + self.dropout2(x)
            t2 = self.norm3(tgt)
            tgt = tgt + self.dropout3(self.linear2(self.dropout(F.relu(self.linear1(t2)))))
            return tgt, attn
        qk = addpos(tgt, query_pos)                            # forward_post, :492-529
        tgt = self.norm1(tgt + self.dropout1(self.self_attn(qk, qk, tgt)))
        x, attn = self.multihead_attn(addpos(tgt, query_pos), addpos(memory, pos), reference_point,
                                      reference_angle, enc_xyz)
        tgt = self.norm2(tgt + self.dropout2(x))
        tgt = self.norm3(tgt + self.dropout3(self.linear2(self.dropout(F.relu(self.linear1(tgt))))))
        return tgt, attn


class OracleDecoder(nn.Module):
    """TransformerDecoder at the reference defaults: mlp_sep, is_bilable, q_content='random',
    focal loss, querypos_mlp (models/vdetr_transformer.py:105-452; build args at
    models/model_vdetr.py:413-447)."""

    def __init__(self, num_layers=8, d_model=256, nhead=4, dim_ff=256, dropout=0.1, mlp_dropout=0.3,
                 num_queries=1024, box_config=None, share_selfattn=False, **rpe_kw):
        super().__init__()
        cfg = box_config or OracleBoxConfig()
        self.cfg = cfg
        self.num_layers, self.num_queries = num_layers, num_queries
        self.first_layer = OracleFFN(d_model, dim_ff, dropout)
        layer = OracleDecoderLayer(d_model, nhead, dim_ff, dropout, share_selfattn=share_selfattn, **rpe_kw)
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(d_model)
        self.query_pos_projection = nn.ModuleList([OracleQueryPos(6, d_model) for _ in range(num_layers)])
        self.query_embed = nn.Embedding(num_queries, d_model)
        for _, p in self.named_parameters():                   # _reset_parameters, :236-241
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

        def headset(ncls):
            return nn.ModuleDict({
                "sem_cls_head": _Head(d_model, ncls, mlp_dropout),
                "center_head": _Head(d_model, 3, mlp_dropout),
                "size_head": _Head(d_model, 3, mlp_dropout),
                "angle_cls_head": _Head(d_model, cfg.num_angle_bin, mlp_dropout),
                "angle_residual_head": _Head(d_model, cfg.num_angle_bin, mlp_dropout)})
        proto = headset(cfg.num_semcls)
        self.mlp_heads = nn.ModuleList([headset(1)] + [copy.deepcopy(proto) for _ in range(num_layers)])
        self.pointcls_heads = _Head(d_model, cfg.num_semcls, mlp_dropout)
        prior = -math.log((1 - 0.01) / 0.01)                   # :161-173
        for hs in self.mlp_heads:
            hs["sem_cls_head"].layers[-1].bias.data.fill_(prior)
            for n in ("center_head", "size_head"):
                nn.init.zeros_(hs[n].layers[-1].weight)
                nn.init.zeros_(hs[n].layers[-1].bias)

    # -- models/vdetr_transformer.py:244-333
    def predict_boxes(self, idx, dims, feats, pre_center_n, pre_size_n):
        x = feats.permute(1, 2, 0)                             # [B,C,nQ]
        heads = self.mlp_heads[idx]
        lo, hi = dims
        scene = (hi - lo).unsqueeze(1)
        logits = heads["sem_cls_head"](x).transpose(1, 2)
        pre_c = pre_center_n * scene + lo.unsqueeze(1)
        pre_s = pre_size_n * scene
        creg = heads["center_head"](x).transpose(1, 2).contiguous()
        center = creg * pre_s + pre_c
        center_n = (center - lo.unsqueeze(1)) / scene
        sreg = heads["size_head"](x).transpose(1, 2).contiguous()
        size = torch.exp(sreg) * pre_s
        size_n = size / scene
        a_logits = heads["angle_cls_head"](x).transpose(1, 2)
        a_res_n = heads["angle_residual_head"](x).transpose(1, 2)
        a_res = a_res_n * (math.pi / a_res_n.shape[-1])
        if a_logits.shape[-1] == 1:                            # BoxProcessor.compute_predicted_angle, :48-71
            angle = (a_logits * 0 + a_res * 0).squeeze(-1).clamp(min=0)
            aprob = angle
            angle0 = angle
        else:
            per = 2 * math.pi / self.cfg.num_angle_bin
            aprob, cls = torch.softmax(a_logits, -1).max(-1)
            angle = per * cls.detach() + a_res.gather(2, cls.detach().unsqueeze(-1)).squeeze(-1)
            angle = torch.where(angle > math.pi, angle - 2 * math.pi, angle)
            angle0 = (a_logits.sum(-1) * 0 + a_res.sum(-1) * 0).clamp(min=0)
        corners = self.cfg.box_parametrization_to_corners(center, size, angle)
        corners0 = self.cfg.box_parametrization_to_corners(center, size, angle0)
        with torch.no_grad():
            objectness = logits.sigmoid().max(-1)[0]
        return {"sem_cls_logits": logits, "center_normalized": center_n.contiguous(),
                "center_unnormalized": center, "size_normalized": size_n, "size_unnormalized": size,
                "angle_logits": a_logits, "angle_prob": aprob, "angle_residual": a_res,
                "angle_residual_normalized": a_res_n, "angle_continuous": angle,
                "objectness_prob": objectness, "sem_cls_prob": logits, "box_corners": corners,
                "box_corners_axis_align": corners0, "pre_box_center_unnormalized": pre_c,
                "center_reg": creg, "pre_box_size_unnormalized": pre_s, "size_reg": sreg}

    # -- models/vdetr_transformer.py:335-452
    def forward(self, memory, enc_xyz, dims, enc_center_n, enc_size_n, return_attn=False):
        out = self.first_layer(memory)
        pred = self.predict_boxes(0, dims, self.norm(out), enc_center_n, enc_size_n)
        preds, attns = [pred], []
        score = pred["objectness_prob"].detach()
        if score.shape[1] >= self.num_queries:
            top = torch.topk(score, self.num_queries, dim=1)[1]
        else:
            top = torch.arange(score.shape[1], device=score.device).unsqueeze(0).repeat(score.shape[0], 1)

        def take(t):
            idx = top.view(top.shape + (1,) * (t.dim() - 2)).expand(top.shape + t.shape[2:])
            return torch.gather(t.detach(), 1, idx)
        ref_pts = camera_to_world_corners(take(pred["box_corners"]))
        ref_center, ref_size = take(pred["center_unnormalized"]), take(pred["size_unnormalized"])
        ref_angle = take(pred["angle_continuous"])
        prop_center_n, prop_size_n = take(pred["center_normalized"]), take(pred["size_normalized"])
        B = memory.shape[1]
        out = self.query_embed.weight.unsqueeze(1).repeat(1, B, 1)
        for i, layer in enumerate(self.layers):
            if i > 0:
                ref_pts = camera_to_world_corners(pred["box_corners"].detach())
                ref_center, ref_size = pred["center_unnormalized"].detach(), pred["size_unnormalized"].detach()
                ref_angle = pred["angle_continuous"].detach()
            qpos = self.query_pos_projection[i](torch.cat([ref_center, ref_size], -1)).permute(2, 0, 1)
            out, attn = layer(out, memory, ref_pts, ref_angle, enc_xyz, qpos)
            pred = self.predict_boxes(i + 1, dims, self.norm(out), prop_center_n, prop_size_n)
            preds.append(pred)
            if return_attn:
                attns.append(attn)
        return {"outputs": preds[-1], "aux_outputs": preds[:-1]}, attns


LOSS_KEYS = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits",
             "angle_residual_normalized")


def synthetic_loss(out, seed=99):
    """fixed random linear functional of every decoder output so that every parameter
    receives gradient (the real criterion needs GT boxes + mmcv)."""
    g = torch.Generator().manual_seed(seed)
    loss = 0
    for d in [out["outputs"]] + list(out["aux_outputs"]):
        for k in LOSS_KEYS:
            r = torch.randn(d[k].shape[1:], generator=g).to(d[k].device)
            loss = loss + (d[k].float() * r).sum()
    return loss
