"""Host-logic test of the drop-in modules WITHOUT a GPU: the fused ops are monkeypatched with dense PyTorch
stand-ins (defined here, in the test), so that everything around the kernels -- parameter names, projections,
Morton reordering of the keys, box decoding, layer wiring -- is checked against the reference's golden vectors."""
import os
import types

import numpy as np
import pytest
import torch

import recipe
from oracle import decoder_torch as odt

G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual_normalized",
        "center_unnormalized", "size_unnormalized", "box_corners")


def _dense_rpe_attention(q, k, v, xyz=None, ref_pts=None, ref_angle=None, tables=None, log_scale=512.0, max_value=4.0,
                         impl=None, impl_bwd=None):
    qh = q.permute(0, 2, 1, 3)
    kh, vh = k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    if tables is not None:
        s = s + odt.rpe_bias_torch(ref_pts, xyz, tables, ref_angle, log_scale, max_value)
    return (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3)


@pytest.fixture()
def patched(monkeypatch):
    from vdetr_b200 import ops
    monkeypatch.setattr(ops, "rpe_attention", _dense_rpe_attention)
    return ops


def _build(L, nq, share):
    from vdetr_b200 import vdetr_transformer as vt
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=share)
    first = vt.FFNLayer(d_model=256, dim_feedforward=256, dropout=0.1)
    layer = vt.GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=0.1, pos_for_key=False, args=args)
    return vt.TransformerDecoder(first, layer, vt.ScanNetBoxConfig(), num_layers=L, decoder_dim=256, mlp_dropout=0.3,
                                 mlp_norm="bn1d", mlp_act="relu", mlp_sep=True, pos_for_key=False, num_queries=nq,
                                 cls_loss="focalloss_0.25", is_bilable=True, q_content="random", return_intermediate=True,
                                 args=args)


@pytest.mark.parametrize("name,seed,B,nK,nq,L,share", [("decoder_eval", 31, 2, 96, 32, 2, False),
                                                       ("decoder_share_eval", 51, 1, 64, 16, 1, True)])
def test_module_wiring_matches_reference_golden(patched, name, seed, B, nK, nq, L, share):
    gold = dict(np.load(os.path.join(G, name + ".npz")))
    dec = _build(L, nq, share)
    sd = dec.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert sorted(shapes) == list(gold["shapes_keys"])
    assert [str(shapes[k]) for k in sorted(shapes)] == list(gold["shapes_vals"])
    for k, v in recipe.fill_state_dict(shapes, seed).items():
        if v is not None:
            sd[k] = torch.from_numpy(v)
    dec.load_state_dict(sd)
    dec.eval()
    c = recipe.decoder_case(seed + 1, B, nK)
    feat, xyz = torch.from_numpy(c["feat"]), torch.from_numpy(c["xyz"])
    with torch.no_grad():
        out, _ = dec(None, feat, xyz, xyz, [torch.from_numpy(c["mins"]), torch.from_numpy(c["maxs"])], query_pos=None,
                     enc_box_predictions={"center_normalized": torch.from_numpy(c["center_normalized"]),
                                          "size_normalized": torch.from_numpy(c["size_normalized"])}, enc_box_features=feat)
    for li, d in enumerate(out["aux_outputs"] + [out["outputs"]]):
        for k in KEYS:
            np.testing.assert_allclose(d[k].numpy(), gold[f"l{li}.{k}"], rtol=2e-4, atol=2e-4, err_msg=f"layer {li} {k}")


def test_morton_order_is_a_permutation():
    from vdetr_b200.vdetr_transformer import morton_order
    xyz = torch.rand(3, 257, 3) * torch.tensor([8.0, 8.0, 3.0])
    perm = morton_order(xyz)
    assert perm.shape == (3, 257)
    assert all(sorted(p.tolist()) == list(range(257)) for p in perm)
