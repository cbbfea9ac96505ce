"""Decoder-level GPU parity: the drop-in TransformerDecoder (vdetr_b200.vdetr_transformer) against
  (1) golden vectors of the UNMODIFIED reference decoder (tests/golden/decoder_*.npz), and
  (2) the CPU oracle port at a larger, ragged size, forward and backward.
Weights come from tests/golden/recipe.py, so all three implementations hold the same state_dict."""
import os
import types

import numpy as np
import pytest
import torch

import recipe
from oracle import decoder_torch as odt

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _strict_fp32():
    """The decoder picks its queries with a top-k over proposal scores: the 1e-3 noise of TF32 convolutions
    (cuDNN's default) is enough to swap near-tied proposals and thereby permute the per-query outputs, so the
    PyTorch layers around the kernels run in true fp32 here."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
G = os.path.join(os.path.dirname(__file__), "golden")
KEYS = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual_normalized",
        "center_unnormalized", "size_unnormalized", "box_corners")


def _log(what, value):
    """measured parity figures -> gpurun_out/parity_r2.jsonl (quoted in DESIGN.md section 5)"""
    import json
    try:
        d = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_r2.jsonl"), "a") as f:
            f.write(json.dumps({"what": what, "max_err_over_max": float(value)}) + "\n")
    except OSError:
        pass


def build_product_decoder(L, nq, dropout=0.1, mlp_dropout=0.3, share=False):
    from vdetr_b200 import vdetr_transformer as vt
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=share)
    first = vt.FFNLayer(d_model=256, dim_feedforward=256, dropout=dropout)
    layer = vt.GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=dropout, pos_for_key=False, args=args)
    return vt.TransformerDecoder(first, layer, vt.ScanNetBoxConfig(), num_layers=L, decoder_dim=256, mlp_dropout=mlp_dropout,
                                 mlp_norm="bn1d", mlp_act="relu", mlp_sep=True, pos_for_key=False, num_queries=nq,
                                 cls_loss="focalloss_0.25", is_bilable=True, q_content="random", return_intermediate=True,
                                 args=args)


def _load(dec, seed):
    sd = dec.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    for k, v in recipe.fill_state_dict(shapes, seed).items():
        if v is not None:
            sd[k] = torch.from_numpy(v)
    dec.load_state_dict(sd)
    return shapes


def _run_product(dec, c, train):
    dev = "cuda"
    feat = torch.from_numpy(c["feat"]).to(dev).requires_grad_(train)
    xyz = torch.from_numpy(c["xyz"]).to(dev)
    dims = [torch.from_numpy(c["mins"]).to(dev), torch.from_numpy(c["maxs"]).to(dev)]
    encp = {"center_normalized": torch.from_numpy(c["center_normalized"]).to(dev),
            "size_normalized": torch.from_numpy(c["size_normalized"]).to(dev)}
    with torch.set_grad_enabled(train):
        out, _ = dec(None, feat, xyz, xyz, dims, query_pos=None, enc_box_predictions=encp, enc_box_features=feat)
    return out, feat


@pytest.mark.parametrize("name,seed,B,nK,nq,L,share", [("decoder_eval", 31, 2, 96, 32, 2, False),
                                                       ("decoder_share_eval", 51, 1, 64, 16, 1, True)])
def test_product_decoder_matches_reference_golden(name, seed, B, nK, nq, L, share):
    gold = dict(np.load(os.path.join(G, name + ".npz")))
    dec = build_product_decoder(L, nq, share=share)
    shapes = _load(dec, seed)
    assert sorted(shapes) == list(gold["shapes_keys"])                      # same state_dict surface as the reference
    assert [str(shapes[k]) for k in sorted(shapes)] == list(gold["shapes_vals"])
    dec = dec.cuda().eval()
    out, _ = _run_product(dec, recipe.decoder_case(seed + 1, B, nK), False)
    report = []
    for li, d in enumerate(out["aux_outputs"] + [out["outputs"]]):
        for k in KEYS:
            want = gold[f"l{li}.{k}"]
            got = d[k].float().cpu().numpy()
            rel = np.abs(got - want).max() / (np.abs(want).max() + 1e-6)
            report.append((rel, li, k))
    _log(f"{name} worst output", max(r[0] for r in report))
    bad = [r for r in report if r[0] > 3e-3]         # 3e-3 of the tensor's max after 2 layers (measured: 1.9e-3 / 4e-4)
    assert not bad, "max |err| / max |ref| per (layer, output): " + ", ".join(f"l{li}.{k}={rel:.2e}" for rel, li, k in report)


def _train_case(monkeypatch, impl_fwd, impl_bwd):
    monkeypatch.setenv("VDETR_B200_IMPL", str(impl_fwd))
    monkeypatch.setenv("VDETR_B200_IMPL_BWD", str(impl_bwd))
    dec = build_product_decoder(2, 32, dropout=0.0, mlp_dropout=0.0)
    _load(dec, 41)
    dec = dec.cuda().train()
    out, feat = _run_product(dec, recipe.decoder_case(42, 2, 96), True)
    loss = odt.synthetic_loss(out)
    loss.backward()
    return dec, feat, loss


def test_backward_kernels_alone_match_reference_gradients(monkeypatch):
    """Forward through the fp32 validation kernels, backward through the product (tcgen05, scaled fp16) kernels:
    isolates the backward path -- its gradients agree with the reference's autograd to 2e-3 (6e-3 for the table
    MLPs, whose gradient passes through the fp16 copy of the tables used by the backward's recompute)."""
    gold = dict(np.load(os.path.join(G, "decoder_train.npz")))
    dec, feat, loss = _train_case(monkeypatch, 1, 0)
    assert abs(loss.item() - float(gold["loss"])) <= 1e-4 * abs(float(gold["loss"])) + 1e-3
    g = feat.grad.cpu().numpy()
    assert np.abs(g - gold["dfeat"]).max() <= 2e-3 * np.abs(gold["dfeat"]).max()
    for n, p in dec.named_parameters():
        key = "grad." + n
        if key in gold and not n.endswith("k.bias"):       # d/d(k.bias) is identically 0 (softmax shift invariance)
            got = p.grad.cpu().numpy()
            got = got[::16] if got.ndim == 2 and got.shape[0] > 64 else got
            assert np.abs(got - gold[key]).max() <= 6e-3 * np.abs(gold[key]).max() + 1e-7, n


def test_product_decoder_train_matches_reference_golden_gradients(monkeypatch):
    """Whole product path.  This golden case is deliberately harsh (random O(1) weights, |logits| up to 20, BatchNorm
    on 64 samples): the 1e-3 forward difference of the fp16 S/PV products moves the point at which the gradient is
    evaluated, and the gradient there differs by ~4 % although the backward kernels themselves are exact to 2e-3
    (previous test).  Tolerance: 8 % of the tensor's max (measured 3-7 %); loss 0.5 % (measured 0.2 %)."""
    gold = dict(np.load(os.path.join(G, "decoder_train.npz")))
    dec, feat, loss = _train_case(monkeypatch, 0, 0)
    assert abs(loss.item() - float(gold["loss"])) <= 5e-3 * abs(float(gold["loss"])) + 1e-2, (loss.item(), float(gold["loss"]))
    g = feat.grad.cpu().numpy()
    _log("decoder_train loss (rel)", abs(loss.item() - float(gold["loss"])) / abs(float(gold["loss"])))
    _log("decoder_train dfeat", np.abs(g - gold["dfeat"]).max() / np.abs(gold["dfeat"]).max())
    assert np.abs(g - gold["dfeat"]).max() <= 8e-2 * np.abs(gold["dfeat"]).max(), np.abs(g - gold["dfeat"]).max() / np.abs(gold["dfeat"]).max()
    for n, p in dec.named_parameters():
        key = "grad." + n
        if key in gold and not n.endswith("k.bias"):
            got = p.grad.cpu().numpy()
            got = got[::16] if got.ndim == 2 and got.shape[0] > 64 else got
            want = gold[key]
            _log("decoder_train grad " + n, np.abs(got - want).max() / (np.abs(want).max() + 1e-12))
            assert np.abs(got - want).max() <= 1e-1 * np.abs(want).max() + 1e-6, n


def test_product_decoder_vs_oracle_port_c1_size():
    """BASELINE config C1 (512 keys x 128 queries x 1 layer) plus a ragged variant, product (GPU) vs oracle (CPU)."""
    for (B, nK, nq, L, seed) in [(1, 512, 128, 1, 7), (2, 333, 70, 2, 9)]:
        torch.manual_seed(0)
        ora = odt.OracleDecoder(num_layers=L, num_queries=nq).eval()
        shapes = {k: tuple(v.shape) for k, v in ora.state_dict().items()}
        vals = recipe.fill_state_dict(shapes, seed)
        sd = ora.state_dict()
        for k, v in vals.items():
            if v is not None:
                sd[k] = torch.from_numpy(v)
        ora.load_state_dict(sd)
        dec = build_product_decoder(L, nq)
        dec.load_state_dict(sd)
        dec = dec.cuda().eval()
        c = recipe.decoder_case(seed + 1, B, nK)
        with torch.no_grad():
            want, _ = ora(torch.from_numpy(c["feat"]), torch.from_numpy(c["xyz"]),
                          [torch.from_numpy(c["mins"]), torch.from_numpy(c["maxs"])],
                          torch.from_numpy(c["center_normalized"]), torch.from_numpy(c["size_normalized"]))
        got, _ = _run_product(dec, c, False)
        for li, (dg, dw) in enumerate(zip(got["aux_outputs"] + [got["outputs"]], want["aux_outputs"] + [want["outputs"]])):
            for k in KEYS:
                w = dw[k].numpy()
                gg = dg[k].float().cpu().numpy()
                _log(f"oracle-port B{B} nK{nK} layer {li} {k}", np.abs(gg - w).max() / (np.abs(w).max() + 1e-6))
                tol = (1e-3 if li <= 1 else 3e-3) * (np.abs(w).max() + 1e-6)      # 1e-3 through one decoder layer (north_star), 3e-3 after two
                assert np.abs(gg - w).max() <= tol, f"B{B} layer {li} {k}: {np.abs(gg - w).max():.3e} > {tol:.3e}"


# ---------------------------------------------------------------------------------------------------------
# GlobalShareCrossAttention module against the reference module's golden outputs: the fused path, and the
# materialising route (return_attn_weights / attn_mask) whose `attn` tensor and mask semantics
# (models/vdetr_transformer.py:743-758: bool mask -> logit := -100, float mask added) must be the reference's.
# ---------------------------------------------------------------------------------------------------------
def _xattn_module(seed):
    from vdetr_b200 import vdetr_transformer as vt
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128)
    mod = vt.GlobalShareCrossAttention(256, 4, args=args)
    sd = mod.state_dict()
    for k, v in recipe.xattn_params(seed).items():
        sd[k] = torch.from_numpy(v)
    mod.load_state_dict(sd)
    return mod.cuda().eval()


def _rel(got, want):
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-12))


def test_cross_attention_module_matches_reference_golden_fused_and_materialised():
    gold = dict(np.load(os.path.join(G, "xattn_small.npz")))
    mod = _xattn_module(11)
    c = recipe.xattn_case(12, 2, 24, 80, False, 0.3)
    t = lambda a: torch.from_numpy(a).cuda()          # noqa: E731
    ref = t(gold["ref_pts"])
    with torch.no_grad():
        x_f, attn_f = mod(t(c["query"]), t(c["key"]), ref, None, t(c["xyz"]))
        x_m, attn_m = mod(t(c["query"]), t(c["key"]), ref, None, t(c["xyz"]), need_weights=True)
    assert attn_f is None
    assert _rel(x_f.cpu().numpy(), gold["x"]) <= 1e-3, _rel(x_f.cpu().numpy(), gold["x"])
    assert _rel(x_m.cpu().numpy(), gold["x"]) <= 1e-4
    assert attn_m.shape == gold["attn"].shape and _rel(attn_m.cpu().numpy(), gold["attn"]) <= 1e-4


def test_cross_attention_attn_mask_semantics_match_reference():
    gold = dict(np.load(os.path.join(G, "xattn_mask.npz")))
    ref_pts = np.load(os.path.join(G, "xattn_small.npz"))["ref_pts"]
    mod = _xattn_module(11)
    c = recipe.xattn_case(12, 2, 24, 80, False, 0.3)
    mb, mf = recipe.xattn_masks(13, 2, 24, 80)
    t = lambda a: torch.from_numpy(a).cuda()          # noqa: E731
    with torch.no_grad():
        for tag, m in (("bool", mb), ("float", mf)):
            x, attn = mod(t(c["query"]), t(c["key"]), t(ref_pts), None, t(c["xyz"]), attn_mask=t(m))
            assert _rel(attn.cpu().numpy(), gold["attn_" + tag]) <= 1e-4, tag
            assert _rel(x.cpu().numpy(), gold["x_" + tag]) <= 1e-4, tag


def test_decoder_layer_return_attn_weights_path():
    """TransformerDecoder(..., return_attn_weights=True) returns the stacked [L,B,H,nQ,nK] probabilities (:447-452) and
    the same predictions as the fused path."""
    dec = build_product_decoder(2, 32)
    _load(dec, 31)
    dec = dec.cuda().eval()
    c = recipe.decoder_case(32, 2, 96)
    dev = "cuda"
    feat = torch.from_numpy(c["feat"]).to(dev)
    xyz = torch.from_numpy(c["xyz"]).to(dev)
    dims = [torch.from_numpy(c["mins"]).to(dev), torch.from_numpy(c["maxs"]).to(dev)]
    encp = {"center_normalized": torch.from_numpy(c["center_normalized"]).to(dev),
            "size_normalized": torch.from_numpy(c["size_normalized"]).to(dev)}
    with torch.no_grad():
        o1, a1 = dec(None, feat, xyz, xyz, dims, enc_box_predictions=encp, enc_box_features=feat)
        o2, a2 = dec(None, feat, xyz, xyz, dims, enc_box_predictions=encp, enc_box_features=feat, return_attn_weights=True)
    assert a2.shape == (2, 2, 4, 32, 96)
    assert torch.allclose(a2.sum(-1), torch.ones_like(a2.sum(-1)), atol=1e-4)
    for k in KEYS:
        w = o2["outputs"][k].float()
        assert (o1["outputs"][k].float() - w).abs().max().item() <= 2e-3 * (w.abs().max().item() + 1e-6), k


def test_fused_box_decode_and_grouped_heads_equal_the_plain_paths():
    """The fused box-decode kernel (csrc/boxdecode.cu) and the grouped evaluation of the 5 heads of a level against the
    op-by-op PyTorch formulation of the same decoder: outputs, corners and the gradient of the encoder features."""
    c = recipe.decoder_case(42, 2, 96)
    results = {}
    for fuse, group in ((True, True), (False, False)):
        # default (xavier) initialisation: the harsh recipe weights of the golden cases make the two-layer decoder chaotic
        # enough to amplify last-bit differences (FMA contraction, GEMM summation order) to 1e-3
        torch.manual_seed(0)
        dec = build_product_decoder(2, 32, dropout=0.0, mlp_dropout=0.0)
        with torch.no_grad():
            for hs in dec.mlp_heads:
                for name in ("center_head", "size_head"):         # zero-initialised in the reference: give the decode something to do
                    hs[name].layers[-1].weight.normal_(0.0, 0.05)
                    hs[name].layers[-1].bias.normal_(0.0, 0.05)
        dec = dec.cuda().train()
        dec.fuse_box_decode, dec.group_heads = fuse, group
        out, feat = _run_product(dec, c, True)
        loss = odt.synthetic_loss(out) + sum(d["box_corners"].square().sum() + d["size_unnormalized"].sum()
                                             for d in out["aux_outputs"] + [out["outputs"]])
        loss.backward()
        results[fuse] = (out, feat.grad.clone(), {n: p.grad.clone() for n, p in dec.named_parameters() if p.grad is not None})
    (oa, ga, pa), (ob, gb, pb) = results[True], results[False]
    for da, db in zip(oa["aux_outputs"] + [oa["outputs"]], ob["aux_outputs"] + [ob["outputs"]]):
        assert "_reference_point_lidar" not in da
        for k in KEYS + ("box_corners_axis_align", "pre_box_center_unnormalized", "pre_box_size_unnormalized"):
            w = db[k].float()
            # (batched vs plain GEMMs sum in different orders, and BatchNorm over 64 samples amplifies the last bits)
            assert (da[k].float() - w).abs().max().item() <= 3e-4 * (w.abs().max().item() + 1e-6), k
    # (two evaluation orders of the same graph: training BatchNorm over 64 samples amplifies their last-bit differences; 2.5e-3
    # measured with the batch-first token order, 1e-3 ... 2e-3 with the sequence-first one)
    assert (ga - gb).abs().max().item() <= 4e-3 * gb.abs().max().item()
    for n in pb:
        if n.endswith("center_head.layers.8.weight") or n.endswith("size_head.layers.8.weight"):
            assert (pa[n] - pb[n]).abs().max().item() <= 4e-3 * (pb[n].abs().max().item() + 1e-9), n


def test_decoder_c2_shape_vs_oracle_port_per_layer(monkeypatch):
    """BASELINE config C2: one ScanNet-shaped scene, 4096 keys x 1024 queries x 8 decoder layers, eval mode -- the product on
    the GPU against the CPU oracle port, error per layer (logged to gpurun_out/parity_r2.jsonl).  Weights: the reference's
    own initialisation (xavier_uniform, zero-initialised centre / size heads) under a fixed seed, shared through the
    state_dict.  The 1024 queries are the top-k of 4096 proposal scores: adjacent scores are ~1e-4 apart, so the 1e-6
    difference between a CPU and a GPU fp32 GEMM swaps a few near-tied ranks and thereby permutes per-query outputs.  The
    numeric comparison therefore runs the product on the oracle's selection, and the product's own selection is checked
    separately (same set up to near-ties)."""
    B, nK, nq, L = 1, 4096, 1024, 8
    torch.manual_seed(0)
    ora = odt.OracleDecoder(num_layers=L, num_queries=nq).eval()
    with torch.no_grad():                                   # non-trivial box refinement: the reference zero-initialises these
        for hs in ora.mlp_heads:
            for name in ("center_head", "size_head"):
                hs[name].layers[-1].weight.normal_(0.0, 0.02)
    dec = build_product_decoder(L, nq)
    dec.load_state_dict(ora.state_dict())
    dec = dec.cuda().eval()
    c = recipe.decoder_case(77, B, nK)
    with torch.no_grad():
        want, _ = ora(torch.from_numpy(c["feat"]), torch.from_numpy(c["xyz"]), [torch.from_numpy(c["mins"]), torch.from_numpy(c["maxs"])],
                      torch.from_numpy(c["center_normalized"]), torch.from_numpy(c["size_normalized"]))
    want_top = torch.topk(want["aux_outputs"][0]["objectness_prob"], nq, dim=1)[1]
    real_topk, seen = torch.topk, {}

    def topk_on_oracle_selection(score, k, dim=-1, **kw):
        vals, idx = real_topk(score, k, dim=dim, **kw)
        if k == nq and tuple(score.shape) == (B, nK):
            seen["own"] = idx.cpu()
            return vals, want_top.to(idx.device)
        return vals, idx
    monkeypatch.setattr(torch, "topk", topk_on_oracle_selection)
    got, _ = _run_product(dec, c, False)
    monkeypatch.setattr(torch, "topk", real_topk)
    assert "own" in seen
    common = len(set(seen["own"][0].tolist()) & set(want_top[0].tolist()))
    _log("C2 top-k selection: fraction of the oracle's 1024 proposals the product selects itself", common / nq)
    assert common >= nq - 8
    worst = []
    for li, (dg, dw) in enumerate(zip(got["aux_outputs"] + [got["outputs"]], want["aux_outputs"] + [want["outputs"]])):
        errs = {k: float(np.abs(dg[k].float().cpu().numpy() - dw[k].numpy()).max() / (np.abs(dw[k].numpy()).max() + 1e-6)) for k in KEYS}
        worst.append(max(errs.values()))
        _log(f"C2 4096x1024x8 eval vs oracle port, level {li} worst of {len(KEYS)} outputs", worst[-1])
    assert worst[0] <= 1e-5                                  # proposal stage: no attention involved
    assert worst[1] <= 1e-3, worst                           # one decoder layer: north_star's tolerance
    assert max(worst) <= 5e-3, worst                         # eight layers deep
