"""Deterministic input/weight recipes shared by make_golden.py (run against the real reference in
the build container) and by the tests (run anywhere).  Only numpy's legacy MT19937 RandomState is
used, whose stream is stable across numpy versions, so fixtures need to store outputs only."""
from __future__ import annotations

import numpy as np

XATTN_PARAM_SHAPES = {
    # name -> shape, for GlobalShareCrossAttention(dim=256, num_heads=4, rpe_dim=128)
    "q.weight": (256, 256), "q.bias": (256,),
    "k.weight": (64, 256), "k.bias": (64,),
    "v.weight": (64, 256), "v.bias": (64,),
    "proj.weight": (256, 256), "proj.bias": (256,),
}
for _i in range(8):
    XATTN_PARAM_SHAPES[f"cpb_mlps.{_i}.0.weight"] = (128, 3)
    XATTN_PARAM_SHAPES[f"cpb_mlps.{_i}.0.bias"] = (128,)
    XATTN_PARAM_SHAPES[f"cpb_mlps.{_i}.2.weight"] = (4, 128)


def xattn_params(seed: int) -> dict:
    rs = np.random.RandomState(seed)
    out = {}
    for name in sorted(XATTN_PARAM_SHAPES):
        shp = XATTN_PARAM_SHAPES[name]
        if name.startswith("cpb_mlps"):
            scale = 0.6 if name.endswith("0.weight") else (0.5 if name.endswith("bias") else 0.04)
        elif name.endswith("weight"):
            scale = 1.0 / np.sqrt(shp[1])
        else:
            scale = 0.1
        out[name] = (rs.standard_normal(shp) * scale).astype(np.float32)
    return out


def xattn_case(seed: int, B: int, nQ: int, nK: int, rotated: bool = False, far: float = 0.0):
    """Synthetic GlobalShareCrossAttention inputs.

    Room 8 x 8 x 3 m on a 0.04 m lattice; a fraction ``far`` of the keys is pushed
    10-25 m away so that grid_sample's zero padding / border blending is exercised.
    """
    rs = np.random.RandomState(seed)
    room = np.array([8.0, 8.0, 3.0], dtype=np.float32)
    xyz = (np.round(rs.rand(B, nK, 3) * room / 0.04) * 0.04).astype(np.float32)
    if far > 0:
        m = rs.rand(B, nK) < far
        xyz[m] += (rs.rand(int(m.sum()), 3).astype(np.float32) * 15 + 10) * \
            np.sign(rs.rand(int(m.sum()), 3).astype(np.float32) - 0.5)
    center = (rs.rand(B, nQ, 3) * room).astype(np.float32)
    size = (rs.rand(B, nQ, 3) + 0.3).astype(np.float32)
    angle = (rs.rand(B, nQ).astype(np.float32) - 0.5) * 2.5 if rotated else None
    query = rs.standard_normal((nQ, B, 256)).astype(np.float32)
    key = rs.standard_normal((nK, B, 256)).astype(np.float32)
    dout = rs.standard_normal((nQ, B, 256)).astype(np.float32)
    return dict(xyz=xyz, center=center, size=size, angle=angle, query=query, key=key, dout=dout)


def xattn_masks(seed: int, B: int, nQ: int, nK: int):
    """attn_mask test inputs [B,nQ,nK]: a bool mask (about 30 % masked, never a whole row) and a float mask."""
    rs = np.random.RandomState(seed)
    mb = rs.rand(B, nQ, nK) < 0.3
    mb[:, :, 0] = False
    mf = (rs.standard_normal((B, nQ, nK)) * 2.0).astype(np.float32)
    return mb, mf


def fill_state_dict(shapes: dict, seed: int, scale: float = 0.08) -> dict:
    """Deterministic values for an arbitrary {name: shape} (sorted-key order), used for whole-decoder
    golden vectors.  BatchNorm running_var gets positive values, num_batches_tracked zeros."""
    rs = np.random.RandomState(seed)
    out = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            out[name] = np.zeros(shp, dtype=np.int64)
        elif name.endswith("running_var"):
            out[name] = (rs.rand(*shp) + 0.5).astype(np.float32)
        elif name.endswith("relative_coords_table"):
            out[name] = None                       # keep the module's own buffer
        elif len(shp) <= 1 and (".norm" in name or name.startswith("norm") or ".1.weight" in name
                                or ".5.weight" in name) and name.endswith("weight"):
            out[name] = (1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif len(shp) <= 1:
            out[name] = (0.05 * rs.standard_normal(shp)).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            out[name] = (rs.standard_normal(shp) / np.sqrt(max(fan_in, 1))).astype(np.float32)
        _ = scale
    return out


def decoder_case(seed: int, B: int, nK: int):
    rs = np.random.RandomState(seed)
    room = np.array([8.0, 8.0, 3.0], dtype=np.float32)
    xyz = (np.round(rs.rand(B, nK, 3) * room / 0.04) * 0.04).astype(np.float32)
    feat = rs.standard_normal((nK, B, 256)).astype(np.float32)
    size = (rs.rand(B, nK, 3) + 0.3).astype(np.float32)
    mins, maxs = xyz.min(1), xyz.max(1)
    sc = maxs - mins
    return dict(xyz=xyz, feat=feat, mins=mins, maxs=maxs,
                center_normalized=((xyz - mins[:, None]) / sc[:, None]).astype(np.float32),
                size_normalized=(size / sc[:, None]).astype(np.float32))
