"""CPU oracle for the Vertex-RPE cross-attention core (numpy, float64/float32).

TEST INFRASTRUCTURE ONLY.  Nothing under ``v-detr_b200/`` may import this file; it is
imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg as
the *checker*, never as the thing measured or shipped.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the unmodified reference module
``GlobalShareCrossAttention`` (/root/reference/models/vdetr_transformer.py:656-758) in this
container and committed its outputs/gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those vectors.

Each function cites the reference lines it restates.  The restatement is deliberately written
from the maths, with an explicit 8-corner trilinear gather instead of
``F.grid_sample``, so that it is an independent statement of the algorithm.
"""
from __future__ import annotations

import numpy as np

NUM_VERTS = 8


def lattice(max_value: float = 4.0, num_points: int = 10, dtype=np.float32) -> np.ndarray:
    """relative_coords_table[0,a,b,c,:] = (lin[a], lin[b], lin[c]).

    Reference: models/vdetr_transformer.py:677-682 (meshgrid default indexing 'ij').
    Returns [N,N,N,3].
    """
    lin = np.linspace(-max_value, max_value, num_points, dtype=np.float32).astype(dtype)
    a, b, c = np.meshgrid(lin, lin, lin, indexing="ij")
    return np.stack([a, b, c], axis=-1)


def build_tables(w1, b1, w2, max_value: float = 4.0, num_points: int = 10) -> np.ndarray:
    """Evaluate the 8 per-vertex MLPs (3->hid->ReLU->H, no bias on the 2nd linear) on the lattice.

    Reference: models/vdetr_transformer.py:695-699 (build_cpb_mlp) and :725.
    w1 [8,hid,3], b1 [8,hid], w2 [8,H,hid]  ->  tables [8, N, N, N, H]  (a=z, b=y, c=x axis order,
    because grid_sample maps x -> last table axis).
    """
    lat = lattice(max_value, num_points, dtype=w1.dtype).reshape(-1, 3)          # [N^3,3]
    out = []
    for i in range(NUM_VERTS):
        hid = np.maximum(lat @ w1[i].T + b1[i], 0)
        out.append((hid @ w2[i].T).reshape(num_points, num_points, num_points, -1))
    return np.stack(out, 0)


def rotate_deltas(d: np.ndarray, angle: np.ndarray) -> np.ndarray:
    """Rotated-box branch (angle_type == 'object_coords').

    Reference: models/vdetr_transformer.py:712-720 with roty_batch_tensor (:761-773).
    Net effect: d <- (c*dx - s*dy, s*dx + c*dy, dz) per query.
    d [B,nQ,nK,3], angle [B,nQ].
    """
    c = np.cos(angle)[:, :, None]
    s = np.sin(angle)[:, :, None]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return np.stack([c * dx - s * dy, s * dx + c * dy, dz], -1)


def signed_log(d: np.ndarray, log_scale: float = 512.0, max_value: float = 4.0) -> np.ndarray:
    """g = sign(d) * log2(|d|*log_scale + 1) / log2(8) / max_value.

    Reference: models/vdetr_transformer.py:722-723.
    """
    return np.sign(d) * np.log2(np.abs(d) * d.dtype.type(log_scale) + d.dtype.type(1.0)) \
        / d.dtype.type(3.0) / d.dtype.type(max_value)


def trilinear_weights(g: np.ndarray, num_points: int = 10):
    """align_corners=False pixel coordinates: p = ((g+1)*N - 1)/2, p0 = floor(p), f = p - p0.

    Reference: F.grid_sample semantics at models/vdetr_transformer.py:727-731
    (mode='bilinear' == trilinear for 5-D input, padding_mode='zeros').
    """
    p = ((g + 1) * num_points - 1) / 2
    p0 = np.floor(p)
    return p0.astype(np.int64), (p - p0)


def rpe_bias(ref_pts, xyz, tables, ref_angle=None, log_scale=512.0, max_value=4.0,
             return_parts=False):
    """rpe[b,h,q,k] = sum_i trilerp(T_i[...,h], g(ref[b,q,i] - xyz[b,k])), zero padding.

    Reference: models/vdetr_transformer.py:708-731.
    ref_pts [B,nQ,8,3], xyz [B,nK,3], tables [8,N,N,N,H]  ->  [B,H,nQ,nK]
    """
    B, nQ = ref_pts.shape[:2]
    nK = xyz.shape[1]
    N = tables.shape[1]
    H = tables.shape[-1]
    dt = tables.dtype
    out = np.zeros((B, nQ, nK, H), dtype=dt)
    parts = []
    for i in range(NUM_VERTS):
        d = ref_pts[:, :, None, i, :] - xyz[:, None, :, :]                     # [B,nQ,nK,3]
        if ref_angle is not None:
            d = rotate_deltas(d, ref_angle)
        g = signed_log(d, log_scale, max_value)
        p0, f = trilinear_weights(g, N)
        # grid x -> last table axis (c), y -> b, z -> a
        ix, iy, iz = p0[..., 0], p0[..., 1], p0[..., 2]
        fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]
        T = tables[i]
        for dz in (0, 1):
            wz = fz if dz else 1 - fz
            za = iz + dz
            for dy in (0, 1):
                wy = fy if dy else 1 - fy
                yb = iy + dy
                for dx in (0, 1):
                    wx = fx if dx else 1 - fx
                    xc = ix + dx
                    ok = (za >= 0) & (za < N) & (yb >= 0) & (yb < N) & (xc >= 0) & (xc < N)
                    w = np.where(ok, wz * wy * wx, 0).astype(dt)
                    v = T[np.clip(za, 0, N - 1), np.clip(yb, 0, N - 1), np.clip(xc, 0, N - 1)]
                    out += w[..., None] * v
                    if return_parts:
                        parts.append((i, np.clip(za, 0, N - 1), np.clip(yb, 0, N - 1),
                                      np.clip(xc, 0, N - 1), w))
    out = np.transpose(out, (0, 3, 1, 2))
    if return_parts:
        return out, parts
    return out


def xattn_core_forward(q, k, v, bias, keep=None, p_drop=0.0):
    """P = softmax_k(q k^T + bias); O = dropout(P) v.   MQA: k, v are shared by all heads.

    Reference: models/vdetr_transformer.py:739-753 (attn_mask=None).  Dropout (:751-752, nn.Dropout on the
    probabilities) is expressed through an explicit keep mask [B,H,nQ,nK] so that it can be compared with a kernel that
    draws the same mask: dropout(P) = P * keep / (1 - p_drop).
    q [B,H,nQ,hd] (already multiplied by hd^-0.5), k,v [B,nK,hd], bias [B,H,nQ,nK]
    ->  O [B,H,nQ,hd], P [B,H,nQ,nK] (before dropout), LSE [B,H,nQ] (natural log).
    """
    s = np.einsum("bhqd,bkd->bhqk", q, k) + bias
    m = s.max(-1, keepdims=True)
    e = np.exp(s - m)
    l = e.sum(-1, keepdims=True)
    p = e / l
    pd = p if keep is None else p * keep / (1.0 - p_drop)
    o = np.einsum("bhqk,bkd->bhqd", pd, v)
    return o, p, (m + np.log(l))[..., 0]


def xattn_core_backward(q, k, v, p, o, do, keep=None, p_drop=0.0):
    """Analytic backward of xattn_core_forward.  Returns dq, dk, dv, dbias (= dS)."""
    scale = None if keep is None else keep / (1.0 - p_drop)
    pd = p if keep is None else p * scale
    dv = np.einsum("bhqk,bhqd->bkd", pd, do)
    dp = np.einsum("bhqd,bkd->bhqk", do, v)
    if keep is not None:
        dp = dp * scale
    delta = (do * o).sum(-1, keepdims=True)
    ds = p * (dp - delta)
    dq = np.einsum("bhqk,bkd->bhqd", ds, k)
    dk = np.einsum("bhqk,bhqd->bkd", ds, q)
    return dq, dk, dv, ds


def rpe_bias_backward_tables(ref_pts, xyz, tables_shape, ds, ref_angle=None, log_scale=512.0,
                             max_value=4.0, dtype=np.float64):
    """dTables[i,a,b,c,h] = sum_{b,q,k} dS[b,h,q,k] * w_i,corner(b,q,k)  (adjoint of rpe_bias).

    Reference: autograd of F.grid_sample w.r.t. its input at models/vdetr_transformer.py:727-731.
    """
    dummy = np.zeros(tables_shape, dtype=dtype)
    _, parts = rpe_bias(ref_pts.astype(dtype), xyz.astype(dtype), dummy,
                        None if ref_angle is None else ref_angle.astype(dtype),
                        log_scale, max_value, return_parts=True)
    dT = np.zeros(tables_shape, dtype=dtype)
    dsT = np.transpose(ds, (0, 2, 3, 1)).astype(dtype)                         # [B,nQ,nK,H]
    for (i, za, yb, xc, w) in parts:
        np.add.at(dT[i], (za.ravel(), yb.ravel(), xc.ravel()),
                  (w[..., None] * dsT).reshape(-1, dsT.shape[-1]))
    return dT


def cross_attention_module_forward(query, key, ref_pts, xyz, params, num_heads=4, ref_angle=None,
                                   log_scale=512.0, max_value=4.0, num_points=10):
    """Whole GlobalShareCrossAttention.forward in numpy (eval mode, no mask).

    Reference: models/vdetr_transformer.py:701-758.
    query [nQ,B,D], key [nK,B,D]; params = dict of numpy arrays keyed like the state_dict:
    q.weight q.bias k.weight k.bias v.weight v.bias proj.weight proj.bias
    cpb_mlps.{i}.0.weight cpb_mlps.{i}.0.bias cpb_mlps.{i}.2.weight
    Returns x [nQ,B,D], attn [B,H,nQ,nK].
    """
    nQ, B, D = query.shape
    hd = D // num_heads
    w1 = np.stack([params[f"cpb_mlps.{i}.0.weight"] for i in range(8)])
    b1 = np.stack([params[f"cpb_mlps.{i}.0.bias"] for i in range(8)])
    w2 = np.stack([params[f"cpb_mlps.{i}.2.weight"] for i in range(8)])
    tables = build_tables(w1, b1, w2, max_value, num_points)
    bias = rpe_bias(ref_pts, xyz, tables, ref_angle, log_scale, max_value)
    qb = np.transpose(query, (1, 0, 2))
    kb = np.transpose(key, (1, 0, 2))
    kk = kb @ params["k.weight"].T + params["k.bias"]
    vv = kb @ params["v.weight"].T + params["v.bias"]
    qq = (qb @ params["q.weight"].T + params["q.bias"]).reshape(B, nQ, num_heads, hd)
    qq = np.transpose(qq, (0, 2, 1, 3)) * query.dtype.type(hd ** -0.5)
    o, p, _ = xattn_core_forward(qq, kk, vv, bias)
    x = np.transpose(o, (0, 2, 1, 3)).reshape(B, nQ, D)
    x = x @ params["proj.weight"].T + params["proj.bias"]
    return np.transpose(x, (1, 0, 2)), p


def box_vertices(center, size):
    """8 vertices of axis-aligned boxes in the reference's vertex order (world frame).

    Reference: utils/box_util.py:319-358 composed with flip_axis_to_camera_tensor (:294-301) and
    convert_corners_camera2lidar (models/vdetr_transformer.py:98-102); SURVEY Appendix A:
    vertex i = center + (sx*l, sy*w, sz*h)/2 with signs
    0:(+,+,-) 1:(+,-,-) 2:(-,-,-) 3:(-,+,-) 4:(+,+,+) 5:(+,-,+) 6:(-,-,+) 7:(-,+,+).
    center,size [...,3] -> [...,8,3]
    """
    sx = np.array([+1, +1, -1, -1, +1, +1, -1, -1], dtype=center.dtype)
    sy = np.array([+1, -1, -1, +1, +1, -1, -1, +1], dtype=center.dtype)
    sz = np.array([-1, -1, -1, -1, +1, +1, +1, +1], dtype=center.dtype)
    sgn = np.stack([sx, sy, sz], -1)                                           # [8,3]
    return center[..., None, :] + sgn * (size[..., None, :] / 2)
