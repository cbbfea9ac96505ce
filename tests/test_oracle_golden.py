"""Pins the CPU oracle (oracle/rpe_attention.py numpy, oracle/decoder_torch.py torch port) to golden
vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import recipe
from oracle import rpe_attention as ora
from oracle import decoder_torch as odt

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return dict(np.load(os.path.join(G, name + ".npz"), allow_pickle=False))


def _xattn_inputs(seed, B, nQ, nK, rotated, far):
    params = recipe.xattn_params(seed)
    case = recipe.xattn_case(seed + 1, B, nQ, nK, rotated, far)
    return params, case


@pytest.mark.parametrize("name,seed,B,nQ,nK,rot,far", [("xattn_small", 11, 2, 24, 80, False, 0.3),
                                                       ("xattn_rot", 21, 1, 16, 48, True, 0.2)])
def test_numpy_oracle_matches_reference_module(name, seed, B, nQ, nK, rot, far):
    gold = _load(name)
    params, case = _xattn_inputs(seed, B, nQ, nK, rot, far)
    # vertex ordering -- axis aligned case reproduces the reference's corners
    if not rot:
        np.testing.assert_allclose(ora.box_vertices(case["center"], case["size"]), gold["ref_pts"],
                                   rtol=0, atol=1e-6)
    w1 = np.stack([params[f"cpb_mlps.{i}.0.weight"] for i in range(8)])
    b1 = np.stack([params[f"cpb_mlps.{i}.0.bias"] for i in range(8)])
    w2 = np.stack([params[f"cpb_mlps.{i}.2.weight"] for i in range(8)])
    np.testing.assert_allclose(ora.build_tables(w1, b1, w2), gold["tables"], rtol=1e-5, atol=1e-5)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    x, attn = ora.cross_attention_module_forward(
        case["query"].astype(np.float64), case["key"].astype(np.float64), gold["ref_pts"].astype(np.float64),
        case["xyz"].astype(np.float64), p64, ref_angle=case["angle"].astype(np.float64) if rot else None)
    np.testing.assert_allclose(attn, gold["attn"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(x, gold["x"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name,seed,B,nQ,nK,rot,far", [("xattn_small", 11, 2, 24, 80, False, 0.3),
                                                       ("xattn_rot", 21, 1, 16, 48, True, 0.2)])
def test_numpy_oracle_backward_matches_reference_autograd(name, seed, B, nQ, nK, rot, far):
    gold = _load(name)
    params, case = _xattn_inputs(seed, B, nQ, nK, rot, far)
    f8 = np.float64
    P = {k: v.astype(f8) for k, v in params.items()}
    tables = gold["tables"].astype(f8)
    ref, xyz = gold["ref_pts"].astype(f8), case["xyz"].astype(f8)
    ang = case["angle"].astype(f8) if rot else None
    bias = ora.rpe_bias(ref, xyz, tables, ang)
    qb = np.transpose(case["query"].astype(f8), (1, 0, 2))
    kb = np.transpose(case["key"].astype(f8), (1, 0, 2))
    kk = kb @ P["k.weight"].T + P["k.bias"]
    vv = kb @ P["v.weight"].T + P["v.bias"]
    qq = np.transpose((qb @ P["q.weight"].T + P["q.bias"]).reshape(B, nQ, 4, 64), (0, 2, 1, 3)) * 0.125
    o, p, _ = ora.xattn_core_forward(qq, kk, vv, bias)
    dx = np.transpose(case["dout"].astype(f8), (1, 0, 2))
    do = np.transpose((dx @ P["proj.weight"]).reshape(B, nQ, 4, 64), (0, 2, 1, 3))
    dq, dk, dv, ds = ora.xattn_core_backward(qq, kk, vv, p, o, do)
    dT = ora.rpe_bias_backward_tables(ref, xyz, tables.shape, ds, ang)
    # chain back to the module inputs / parameters and compare with reference autograd
    dqlin = np.transpose(dq, (0, 2, 1, 3)).reshape(B, nQ, 256) * 0.125
    dquery = np.transpose(dqlin @ P["q.weight"], (1, 0, 2))
    dkey = np.transpose(dk @ P["k.weight"] + dv @ P["v.weight"], (1, 0, 2))
    np.testing.assert_allclose(dquery, gold["dquery"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(dkey, gold["dkey"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(dk.sum((0, 1)), gold["grad.k.bias"], rtol=2e-3, atol=2e-5)
    # dTables -> cpb_mlps.{i}.2.weight grad:  T_i = hid_i @ W2_i^T
    lat = ora.lattice().reshape(-1, 3).astype(f8)
    for i in range(8):
        hid = np.maximum(lat @ P[f"cpb_mlps.{i}.0.weight"].T + P[f"cpb_mlps.{i}.0.bias"], 0)
        gw2 = dT[i].reshape(-1, 4).T @ hid
        np.testing.assert_allclose(gw2, gold[f"grad.cpb_mlps.{i}.2.weight"], rtol=3e-3, atol=3e-5)


def _oracle_decoder(L, nq, train, share=False):
    torch.manual_seed(0)
    dec = odt.OracleDecoder(num_layers=L, num_queries=nq, dropout=0.0 if train else 0.1,
                            mlp_dropout=0.0 if train else 0.3, share_selfattn=share)
    return dec


@pytest.mark.parametrize("name,seed,B,nK,nq,L,train,share", [
    ("decoder_eval", 31, 2, 96, 32, 2, False, False),
    ("decoder_train", 41, 2, 96, 32, 2, True, False),
    ("decoder_share_eval", 51, 1, 64, 16, 1, False, True)])
def test_torch_port_matches_reference_decoder(name, seed, B, nK, nq, L, train, share):
    gold = _load(name)
    dec = _oracle_decoder(L, nq, train, share)
    sd = dec.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    # identical state_dict surface (names + shapes) as the reference module
    assert sorted(shapes) == list(gold["shapes_keys"])
    assert [str(shapes[k]) for k in sorted(shapes)] == list(gold["shapes_vals"])
    vals = recipe.fill_state_dict(shapes, seed)
    for k, v in vals.items():
        if v is not None:
            sd[k] = torch.from_numpy(v)
    dec.load_state_dict(sd)
    dec.train(train)
    c = recipe.decoder_case(seed + 1, B, nK)
    feat = torch.from_numpy(c["feat"]).requires_grad_(train)
    with torch.set_grad_enabled(train):
        out, _ = dec(feat, torch.from_numpy(c["xyz"]), [torch.from_numpy(c["mins"]), torch.from_numpy(c["maxs"])],
                     torch.from_numpy(c["center_normalized"]), torch.from_numpy(c["size_normalized"]))
    preds = out["aux_outputs"] + [out["outputs"]]
    for li, d in enumerate(preds):
        for k in ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits",
                  "angle_residual_normalized", "center_unnormalized", "size_unnormalized", "box_corners"):
            np.testing.assert_allclose(d[k].detach().numpy(), gold[f"l{li}.{k}"], rtol=2e-4, atol=2e-4,
                                       err_msg=f"layer {li} {k}")
    if train:
        loss = odt.synthetic_loss(out)
        loss.backward()
        assert abs(loss.item() - float(gold["loss"])) <= 2e-4 * abs(float(gold["loss"])) + 1e-3
        np.testing.assert_allclose(feat.grad.numpy(), gold["dfeat"], rtol=5e-3, atol=5e-4)
        for n, p in dec.named_parameters():
            key = "grad." + n
            if key in gold:
                g = p.grad.numpy()
                g = g[::16] if g.ndim == 2 and g.shape[0] > 64 else g
                np.testing.assert_allclose(g, gold[key], rtol=5e-3, atol=5e-4, err_msg=n)


def test_grid_sample_variant_agrees_with_explicit_gather():
    torch.manual_seed(3)
    tables = torch.randn(8, 10, 10, 10, 4)
    c = recipe.xattn_case(77, 2, 9, 31, True, 0.3)
    ref = torch.from_numpy(ora.box_vertices(c["center"], c["size"]))
    xyz, ang = torch.from_numpy(c["xyz"]), torch.from_numpy(c["angle"])
    for a in (None, ang):
        x = odt.rpe_bias_torch(ref, xyz, tables, a)
        y = odt.rpe_bias_grid_sample(ref, xyz, tables, a)
        torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-5)
