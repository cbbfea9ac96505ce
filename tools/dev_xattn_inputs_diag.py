import os, sys, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import recipe
from vdetr_b200 import ops
from oracle import decoder_torch as odt
from test_decoder_gpu import build_product_decoder, _load, _run_product
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
calls = []
orig = ops.rpe_attention
def rec(q, k, v, xyz=None, ref_pts=None, ref_angle=None, tables=None, *a, **kw):
    calls.append([t.detach().clone() if t is not None else None for t in (q, k, v, xyz, ref_pts, tables)])
    return orig(q, k, v, xyz, ref_pts, ref_angle, tables, *a, **kw)
ops.rpe_attention = rec
dec = build_product_decoder(2, 32, dropout=0.0, mlp_dropout=0.0); _load(dec, 41); dec = dec.cuda().train()
out, feat = _run_product(dec, recipe.decoder_case(42, 2, 96), True)
ops.rpe_attention = orig
torch.manual_seed(0)
for ci, (q, k, v, xyz, ref, tab) in enumerate(calls):
    has_bias = tab is not None
    kk = k.expand(-1, -1, 4, -1) if k.shape[2] == 1 else k
    S = torch.einsum("bqhd,bkhd->bhqk", q, kk)
    if has_bias:
        bias = ops.rpe_bias(xyz, ref, tab); S = S + bias
    P = torch.softmax(S, -1)
    print(f"call {ci} bias={has_bias} q{tuple(q.shape)} k{tuple(k.shape)} |q|max {q.abs().max():.2f} |k|max {k.abs().max():.2f} S std {S.std():.2f} S absmax {S.abs().max():.2f} Pmax mean {P.max(-1)[0].mean():.3f}" + (f" bias absmax {bias.abs().max():.2f} tab absmax {tab.abs().max():.2f}" if has_bias else ""))
    do = torch.randn_like(q)
    res = {}
    for impl in (0, 1):
        qq, k2, v2 = (t.clone().requires_grad_(True) for t in (q, k, v))
        tt = tab.clone().requires_grad_(True) if has_bias else None
        o = orig(qq, k2, v2, xyz, ref, None, tt, impl=impl, impl_bwd=impl)
        o.backward(do)
        res[impl] = (o.detach(), qq.grad, k2.grad, v2.grad, tt.grad if has_bias else None)
    names = ["out", "dq", "dk", "dv", "dT"]
    print("   tc vs simt rel(max):", {n: float((a - b).abs().max() / b.abs().max()) for n, a, b in zip(names, res[0], res[1]) if a is not None})
