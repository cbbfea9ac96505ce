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
                         impl=None, impl_bwd=None, dropout_p=0.0, dropout_seed=None):
    assert dropout_p == 0.0          # eval-mode wiring test
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


def test_token_major_stacks_equal_conv_stacks():
    """helpers.pointwise_tokens (GEMMs on [tokens, C] rows) against the reference formulation it replaces (Conv1d(k=1) /
    BatchNorm1d on [B, C, N], models/helpers.py:17-33, 74-141): same outputs, gradients and running statistics, in
    train and eval mode."""
    import copy
    import torch
    from vdetr_b200 import helpers
    torch.manual_seed(0)
    B, N, C = 3, 50, 256
    mlp = helpers.GenericMLP(input_dim=C, hidden_dims=[C, C], output_dim=18, norm_fn_name="bn1d", activation="relu",
                             use_conv=True, dropout=0.0)
    assert mlp.supports_tokens
    pe = helpers.PositionEmbeddingLearned(6, C)
    for train in (True, False):
        for mod, xin in ((mlp, torch.randn(N, B, C)), (pe, torch.randn(B, N, 6))):
            a, b = copy.deepcopy(mod).train(train), copy.deepcopy(mod).train(train)
            xa, xb = xin.clone().requires_grad_(True), xin.clone().requires_grad_(True)
            if mod is mlp:
                ya = a(xa.permute(1, 2, 0)).transpose(1, 2)                                  # [B, N, out] as the decoder builds it
                yb = b.forward_tokens(xb.reshape(N * B, C)).view(N, B, -1).transpose(0, 1)
            else:
                ya = a(xa).permute(2, 0, 1)                                                   # [N, B, C]
                yb = b.forward_tokens(xb)
            assert torch.allclose(ya, yb, atol=2e-5, rtol=1e-4)
            g = torch.randn_like(ya)
            ya.backward(g)
            yb.backward(g)
            assert torch.allclose(xa.grad, xb.grad, atol=2e-5, rtol=1e-3)
            for (n1, p1), (_, p2) in zip(a.named_parameters(), b.named_parameters()):
                assert torch.allclose(p1.grad, p2.grad, atol=1e-4, rtol=1e-3), n1
            for (n1, b1), (_, b2) in zip(a.named_buffers(), b.named_buffers()):
                assert torch.allclose(b1.float(), b2.float(), atol=1e-5, rtol=1e-4), n1


def test_flat_grad_gather_matches_accumulation():
    """parallel.FlatGradAllReduce: detached gradients + one multi-tensor copy give the flat buffer the same content as
    accumulating into its views, parameters without gradient read as zero, and gather_ is idempotent."""
    import torch
    from vdetr_b200 import parallel
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.ones(5))
    params = list(net.parameters()) + [unused]
    fg = parallel.FlatGradAllReduce(params)
    x = torch.randn(10, 8)
    fg.zero_()
    net(x).square().sum().backward()
    want = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = fg.gather_()
    assert torch.equal(flat, torch.cat([w.reshape(-1) for w in want]))
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(fg.params, fg.views))
    assert torch.equal(fg.gather_(), flat) and torch.equal(fg.sync_(), flat)


def test_grouped_heads_equal_head_by_head():
    """TransformerDecoder._run_heads_grouped (one GEMM / batched GEMMs / grouped BatchNorm for the 5 heads of a level)
    against the head-by-head token path, eval and train mode, including the running-statistics update."""
    import copy
    from vdetr_b200 import vdetr_transformer as vt
    dec = _build(1, 16, False)
    heads = dec.mlp_heads[1]
    g = torch.Generator().manual_seed(3)
    for n in dec.HEAD_NAMES:
        for m in heads[n].modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(generator=g); m.running_var.uniform_(0.5, 2.0, generator=g)
                m.weight.data.normal_(generator=g); m.bias.data.normal_(generator=g)
        heads[n].layers[-1].weight.data.normal_(generator=g); heads[n].layers[-1].bias.data.normal_(generator=g)
        for m in heads[n].modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    x = torch.randn(16, 3, 256, generator=g)
    for train in (False, True):
        dec.train(train)
        stacks = dec._grouped_plan(heads)
        assert stacks is not None
        twin = copy.deepcopy(heads)
        a = dec._run_heads_grouped(stacks, x.reshape(48, 256), 16, 3)
        dec.group_heads = False
        b = dec._run_heads(twin, x)
        dec.group_heads = True
        for n in dec.HEAD_NAMES:
            assert a[n].shape == b[n].shape
            assert (a[n] - b[n]).abs().max().item() <= 1e-4 * (b[n].abs().max().item() + 1.0), n
        if train:
            # gradients through the grouped stack (incl. the _PadStack / _SplitHeadOutputs nodes) equal the head-by-head ones
            w = {n: torch.randn(a[n].shape, generator=g) for n in dec.HEAD_NAMES if n != "angle_cls_head"}   # one head unused
            pa = [p for n in dec.HEAD_NAMES for p in heads[n].parameters()]
            pb = [p for n in dec.HEAD_NAMES for p in twin[n].parameters()]
            ga = torch.autograd.grad(sum((a[n] * w[n]).sum() for n in w), pa, allow_unused=True)
            gb = torch.autograd.grad(sum((b[n] * w[n]).sum() for n in w), pb, allow_unused=True)
            for x1, x2, p_ in zip(ga, gb, pa):
                if x2 is None:
                    assert x1 is None or float(x1.abs().max()) == 0.0
                    continue
                assert x1 is not None and x1.shape == p_.shape
                assert (x1 - x2).abs().max().item() <= 2e-4 * (x2.abs().max().item() + 1e-6)
        for n in dec.HEAD_NAMES:
            for m1, m2 in zip(heads[n].modules(), twin[n].modules()):
                if isinstance(m1, torch.nn.BatchNorm1d):
                    assert torch.allclose(m1.running_mean, m2.running_mean, atol=1e-5)
                    assert torch.allclose(m1.running_var, m2.running_var, atol=1e-5)
                    assert int(m1.num_batches_tracked) == int(m2.num_batches_tracked)
    _ = vt


def test_batched_token_linear_gradients_equal_bmm_autograd():
    """ops._BatchedTokenLinear (dW in the parameter layout, split-K over 512-token slices) against torch.baddbmm's autograd,
    for a contiguous [G, T, in] input and for the [T, G, in] -> transpose view the grouped heads feed into layer 2."""
    from vdetr_b200 import ops
    g = torch.Generator().manual_seed(5)
    G, T, I, O = 3, 2048, 32, 24
    for strided in (False, True):
        base = torch.randn((T, G, I) if strided else (G, T, I), generator=g)
        w = torch.randn(G, O, I, generator=g, requires_grad=True)
        b = torch.randn(G, O, generator=g, requires_grad=True)
        dy = torch.randn(G, T, O, generator=g)
        xa = base.clone().requires_grad_(True)
        xb = base.clone().requires_grad_(True)
        va = xa.transpose(0, 1) if strided else xa
        vb = xb.transpose(0, 1) if strided else xb
        want = torch.baddbmm(b.unsqueeze(1), va, w.transpose(1, 2))
        gw = torch.autograd.grad(want, (xa, w, b), dy)
        got = ops._BatchedTokenLinear.apply(vb, w, b)
        gg = torch.autograd.grad(got, (xb, w, b), dy)
        assert torch.allclose(got, want, atol=1e-5)
        for a_, r_ in zip(gg, gw):
            assert a_.shape == r_.shape and (a_ - r_).abs().max().item() <= 1e-4 * r_.abs().max().item()
        assert gg[1].is_contiguous()
        got2 = ops._BatchedTokenLinear.apply(vb, w, None)
        assert torch.allclose(got2, torch.bmm(va, w.transpose(1, 2)), atol=1e-5)


def test_batch_first_backed_tokens_give_the_same_heads_and_gradients():
    """The decoder keeps token activations in batch-first memory behind [nQ, B, C] views (ops.batch_first_backed): the grouped
    heads on tokens in (b, q) order, incl. the _SplitHeadOutputs backward, against the sequence-first evaluation; and the
    layout-preserving wrappers of ops.linear."""
    import copy
    from vdetr_b200 import ops
    dec = _build(1, 16, False).train()
    heads = dec.mlp_heads[1]
    g = torch.Generator().manual_seed(11)
    for n in dec.HEAD_NAMES:
        heads[n].layers[-1].weight.data.normal_(generator=g); heads[n].layers[-1].bias.data.normal_(generator=g)
        for m in heads[n].modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    nQ, B, C = 16, 3, 256
    x = torch.randn(nQ, B, C, generator=g)
    xa = x.clone().requires_grad_(True)
    xb = x.transpose(0, 1).contiguous().requires_grad_(True)          # batch-first memory
    vb = xb.transpose(0, 1)                                           # the [nQ, B, C] view the decoder hands on
    assert ops.batch_first_backed(vb) and not ops.batch_first_backed(xa)
    twin = copy.deepcopy(heads)
    a = dec._run_heads_grouped(dec._grouped_plan(heads), xa.reshape(nQ * B, C), nQ, B, False)
    b = dec._run_heads_grouped(dec._grouped_plan(twin), vb.transpose(0, 1).reshape(B * nQ, C), nQ, B, True)
    w = {n: torch.randn(a[n].shape, generator=g) for n in dec.HEAD_NAMES}
    ga, = torch.autograd.grad(sum((a[n] * w[n]).sum() for n in w), xa)
    gb, = torch.autograd.grad(sum((b[n] * w[n]).sum() for n in w), xb)
    for n in dec.HEAD_NAMES:
        assert a[n].shape == b[n].shape == (B, nQ, a[n].shape[-1])
        assert (a[n] - b[n]).abs().max().item() <= 1e-4 * (a[n].abs().max().item() + 1.0), n
    assert (ga - gb.transpose(0, 1)).abs().max().item() <= 2e-4 * ga.abs().max().item()
    # ops.linear on a batch-first-backed view: same values, same kind of view
    lin = torch.nn.Linear(C, 64)
    y = ops.linear(vb, lin.weight, lin.bias)
    assert ops.batch_first_backed(y) and torch.allclose(y, torch.nn.functional.linear(x, lin.weight, lin.bias), atol=1e-5)
