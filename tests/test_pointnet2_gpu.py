"""GPU parity of the pointnet2 ops: CUDA kernels (through the C ABI) vs the C oracle, and -- when the rebuilt
reference extension oracle/_ref/pn2_ref_ext.so is present -- vs the reference kernels themselves."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import pn2_util as U

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pu():
    import vdetr_b200.pointnet2_utils as pu
    return pu


def _ref_ext():
    so = os.path.join(ROOT, "oracle", "_ref", "pn2_ref_ext.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("pn2_ref_ext", so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("b,n,m,zero", [(2, 700, 64, 0.05), (1, 5, 9, 0.0), (3, 4096, 512, 0.0), (2, 9000, 300, 0.02),
                                        (1, 20000, 1024, 0.0)])
def test_fps_bit_exact_vs_c_oracle(b, n, m, zero):
    pts = U.lattice_cloud(n + m, b, n, zero)
    if n < 10000:
        pts = (np.round(pts / 0.2) * 0.2).astype(np.float32)          # force ties
    want = U.ref_fps(pts, m)
    got = _pu().furthest_point_sample(torch.from_numpy(pts).cuda(), m).cpu().numpy()
    assert got.dtype == np.int32 and (got == want).all()


def test_fps_scannet_size_vs_reference_ext():
    ext = _ref_ext()
    pts = torch.from_numpy(U.lattice_cloud(7, 2, 50000)).cuda()
    got = _pu().furthest_point_sample(pts, 4096)
    if ext is not None:
        want = ext.furthest_point_sampling(pts, 4096)
        assert (got == want).all()
    # size-independent properties: starts at 0, all picks distinct, greedy max-min at a probe step
    g = got.cpu().numpy()
    assert (g[:, 0] == 0).all() and all(len(set(r.tolist())) == 4096 for r in g)
    p = pts[0].double()
    sel = p[got[0, :1000].long()]
    d = torch.cdist(p, sel).min(1)[0]
    assert abs(d[got[0, 1000]].item() - d.max().item()) < 1e-6


def test_reference_ext_agrees_with_c_oracle():
    """Pins oracle/pointnet2_ref.c against the UNMODIFIED reference kernels rebuilt for sm_100a."""
    ext = _ref_ext()
    if ext is None:
        pytest.skip("oracle/_ref/pn2_ref_ext.so not built")
    pts = (np.round(U.lattice_cloud(3, 2, 3000, 0.03) / 0.2) * 0.2).astype(np.float32)
    t = torch.from_numpy(pts).cuda()
    assert (ext.furthest_point_sampling(t, 200).cpu().numpy() == U.ref_fps(pts, 200)).all()
    ctr = t[:, :300].contiguous()
    assert (ext.ball_query(ctr, t, 0.45, 16).cpu().numpy() == U.ref_ball_query(pts[:, :300], pts, 0.45, 16)).all()


@pytest.mark.parametrize("n,m,r,ns", [(500, 40, 0.2, 16), (20000, 2048, 0.2, 64), (3000, 77, 0.45, 5)])
def test_ball_query_bit_exact(n, m, r, ns):
    pts = U.lattice_cloud(n, 2, n)
    ctr = pts[:, :m].copy()
    ctr[:, -1] += 100.0                                            # one centre with no neighbour
    want = U.ref_ball_query(ctr, pts, r, ns)
    got = _pu().ball_query(r, ns, torch.from_numpy(pts).cuda(), torch.from_numpy(ctr).cuda()).cpu().numpy()
    assert (got == want).all()
    ext = _ref_ext()
    if ext is not None:
        ref = ext.ball_query(torch.from_numpy(ctr).cuda(), torch.from_numpy(pts).cuda(), r, ns).cpu().numpy()
        assert (got == ref).all()


def test_gather_and_group_forward_backward():
    pu = _pu()
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(2, 37, 900, generator=g).cuda().requires_grad_(True)
    idx = torch.stack([torch.randperm(900, generator=g)[:128] for _ in range(2)]).int().cuda()
    out = pu.gather_operation(feats, idx)
    want = torch.gather(feats, 2, idx.long().unsqueeze(1).expand(-1, 37, -1))
    assert torch.equal(out, want)
    go = torch.randn_like(out)
    out.backward(go)
    gw = torch.zeros_like(feats).scatter_add_(2, idx.long().unsqueeze(1).expand(-1, 37, -1), go)
    assert torch.equal(feats.grad, gw)                             # unique indices -> exact
    gidx = torch.randint(0, 900, (2, 50, 8), generator=g).int().cuda()
    f2 = feats.detach().clone().requires_grad_(True)
    grp = pu.grouping_operation(f2, gidx)
    wantg = torch.gather(f2.unsqueeze(2).expand(-1, -1, 50, -1), 3, gidx.long().unsqueeze(1).expand(-1, 37, -1, -1))
    assert torch.equal(grp, wantg)
    gg = torch.randn_like(grp)
    grp.backward(gg)
    ref = torch.zeros(2, 37, 900, device="cuda").scatter_add_(2, gidx.long().view(2, 1, -1).expand(-1, 37, -1), gg.view(2, 37, -1))
    torch.testing.assert_close(f2.grad, ref, rtol=1e-5, atol=1e-5)


def test_cpu_and_dtype_inputs_are_rejected():
    pu = _pu()
    with pytest.raises(RuntimeError):
        pu.furthest_point_sample(torch.zeros(1, 8, 3), 4)
    with pytest.raises(RuntimeError):
        pu.gather_operation(torch.zeros(1, 4, 8).cuda(), torch.zeros(1, 4).long().cuda())
