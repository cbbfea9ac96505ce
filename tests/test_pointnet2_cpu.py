"""CPU tests of the pointnet2 oracle (oracle/pointnet2_ref.c) and of the tie-break theory the CUDA FPS
kernel relies on (slot order == the reference's strided scan + shared-memory tree)."""
import numpy as np
import pytest

import pn2_util as U


@pytest.mark.parametrize("n,m,seed", [(700, 64, 1), (512, 40, 2), (1300, 100, 3), (37, 20, 4), (5, 9, 5), (2048, 128, 6)])
def test_slot_order_equals_literal_reference_emulation(n, m, seed):
    pts = U.lattice_cloud(seed, 2, n, zero_frac=0.05)
    # coarse lattice -> many exact ties
    pts = np.round(pts / 0.5) * 0.5
    a = U.ref_fps(pts.astype(np.float32), m)
    b = U.fps_slot_order_numpy(pts.astype(np.float32), m)
    assert (a == b).all()


def test_fps_all_points_skipped_returns_zeros():
    pts = np.full((1, 64, 3), 0.01, dtype=np.float32)
    assert (U.ref_fps(pts, 8) == 0).all()


def test_fps_properties():
    pts = U.lattice_cloud(9, 1, 4000)
    idx = U.ref_fps(pts, 256)[0]
    assert idx[0] == 0 and len(set(idx.tolist())) == 256          # distinct while m <= #distinct points
    # greedy property: each pick maximises the distance to the already selected set
    sel = pts[0][idx[:50]]
    d = ((pts[0][:, None, :] - sel[None]) ** 2).sum(-1).min(1)
    assert abs(d[idx[50]] - d.max()) <= 1e-5 * d.max()


def test_ball_query_semantics():
    rs = np.random.RandomState(0)
    xyz = rs.rand(2, 500, 3).astype(np.float32)
    ctr = xyz[:, :40].copy()
    idx = U.ref_ball_query(ctr, xyz, 0.2, 16)
    d2 = ((ctr[:, :, None, :].astype(np.float64) - xyz[:, None].astype(np.float64)) ** 2).sum(-1)
    for b in range(2):
        for j in range(40):
            hits = np.nonzero(d2[b, j] < 0.2 ** 2 - 1e-9)[0]
            row = idx[b, j]
            k = min(len(hits), 16)
            assert (row[:k] == hits[:k]).all()
            assert (row[k:] == (hits[0] if len(hits) else 0)).all()   # padded with the first hit
    far = U.ref_ball_query(ctr + 100, xyz, 0.2, 4)
    assert (far == 0).all()
