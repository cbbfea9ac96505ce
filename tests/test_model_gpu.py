"""GPU test of the model wrapper downstream of the (out-of-scope) sparse backbone: batched FPS + gather on a
synthetic sparse-tensor output, projection, proposal anchors and the decoder (ModelVDETR.forward_from_backbone)."""
import types

import numpy as np
import pytest
import torch

import pn2_util as U

pytestmark = pytest.mark.gpu


def _args():
    return types.SimpleNamespace(
        dec_dim=256, dec_ffn_dim=256, dec_dropout=0.1, dec_nhead=4, dec_nlayers=3, pos_for_key=False, mlp_dropout=0.3,
        mlp_norm="bn1d", mlp_act="relu", mlp_sep=True, nqueries=64, cls_loss="focalloss_0.25", is_bilable=True,
        q_content="random", log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=False,
        enc_dim=256, querypos_mlp=True, minkowski=True, inplanes=64, num_stages=4, voxel_size=0.01, preenc_npoints=256,
        use_fpn=True, layer_idx=0, proj_nohid=True, woexpand_conv=False, depth=34, use_color=False, xyz_color=False)


def test_forward_from_backbone_and_batched_fps():
    from vdetr_b200 import model_vdetr as mv
    from vdetr_b200.vdetr_transformer import ScanNetBoxConfig
    torch.manual_seed(0)
    model = mv.build_vdetr(_args(), ScanNetBoxConfig(), pre_encoder=None).cuda().eval()
    B, N = 2, 3000
    pts = U.lattice_cloud(3, B, N)
    vox = np.round(pts / 0.01).astype(np.int32)
    coords = torch.from_numpy(np.concatenate([np.concatenate([np.full((N, 1), b, np.int32), vox[b]], 1) for b in range(B)])).cuda()
    feats = torch.randn(B * N, 256).cuda()
    xyz, f, inds = model.sample_backbone_output(coords, feats, 256)
    want = U.ref_fps((vox.astype(np.float32) * 0.01).astype(np.float32), 256)
    assert (inds.cpu().numpy() == want).all()                                   # batched launch == per-scene oracle
    assert xyz.shape == (B, 256, 3) and f.shape == (B, 256, 256)
    assert torch.equal(f, torch.gather(feats.view(B, N, 256).transpose(1, 2), 2, inds.long().unsqueeze(1).expand(-1, 256, -1)))
    dims = [xyz.min(1)[0], xyz.max(1)[0]]
    with torch.no_grad():
        out = model.forward_from_backbone(xyz, f.permute(2, 0, 1), inds, dims)
    assert set(["outputs", "aux_outputs", "seed_inds", "seed_xyz", "enc_outputs"]) <= set(out)
    assert out["outputs"]["sem_cls_logits"].shape == (B, 64, 18) and len(out["aux_outputs"]) == 2
    assert all(torch.isfinite(v).all() for v in out["outputs"].values() if torch.is_tensor(v))


def test_ragged_fps_matches_per_scene_oracle():
    """Scenes with different voxel counts, rows of the sparse tensor interleaved across scenes: one ragged FPS launch (offsets
    on the device) must give the indices of a per-scene B = 1 call (C oracle), in the row order torch.where would produce."""
    from vdetr_b200 import model_vdetr as mv
    from vdetr_b200.vdetr_transformer import ScanNetBoxConfig
    torch.manual_seed(0)
    model = mv.build_vdetr(_args(), ScanNetBoxConfig(), pre_encoder=None).cuda().eval()
    counts = [3000, 1777, 5003, 300]
    rs = np.random.RandomState(5)
    scenes = [np.round(U.lattice_cloud(10 + i, 1, n)[0] / 0.01).astype(np.int32) for i, n in enumerate(counts)]
    rows = np.concatenate([np.concatenate([np.full((n, 1), b, np.int32), v], 1) for b, (n, v) in enumerate(zip(counts, scenes))])
    perm = rs.permutation(len(rows))                       # interleave the scenes
    rows = rows[perm]
    feats = torch.randn(len(rows), 256).cuda()
    coords = torch.from_numpy(rows).cuda()
    for cap in (None, 8192):
        model.max_points_per_scene = cap
        xyz, f, inds = model.sample_backbone_output(coords, feats, 256, batch_size=len(counts))
        assert xyz.shape == (4, 256, 3) and f.shape == (4, 256, 256)
        for b, n in enumerate(counts):
            sel = np.nonzero(rows[:, 0] == b)[0]                                   # torch.where order
            pts = (rows[sel, 1:].astype(np.float32) * np.float32(0.01)).astype(np.float32)
            want = U.ref_fps(pts[None], 256)[0]
            assert (inds[b].cpu().numpy() == want).all(), (cap, b)
            assert np.array_equal(xyz[b].cpu().numpy(), pts[want])
            assert torch.equal(f[b], feats[torch.from_numpy(sel[want]).cuda()].t())
    model.max_points_per_scene = 1024                      # a wrong bound is reported, not silently sampled
    _, _, inds = model.sample_backbone_output(coords, feats, 256, batch_size=len(counts))
    assert (inds[0] == -1).all() and (inds[3] >= 0).all()
