"""Developer script: the dense-UMMA dTables kernel (VDETR_DT_IMPL=6) against dt3 on identical inputs, then timing at B x 1024 x 4096."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdetr_b200 import ops
from vdetr_b200.vdetr_transformer import morton_order
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sgn = torch.tensor([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1]], dtype=torch.float32).cuda()


def case(B, nQ, nK, seed):
    g = torch.Generator().manual_seed(seed)
    xyz = ((torch.rand(B, nK, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04).cuda()
    perm = morton_order(xyz)
    xyz = torch.gather(xyz, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    center = (torch.rand(B, nQ, 3, generator=g) * torch.tensor([8., 8., 3.])).cuda()
    size = (torch.rand(B, nQ, 3, generator=g) + 0.3).cuda()
    ref = (center[:, :, None, :] + sgn * size[:, :, None, :] / 2).contiguous()
    ds = (torch.randn(B, 4, nQ, nK, generator=g) * torch.exp(torch.randn(B, 4, nQ, nK, generator=g))).cuda()
    return xyz, ref, ds


t0 = torch.zeros(8, 10, 10, 10, 4, device="cuda")
for (b, nq, nk) in ((1, 5, 64), (1, 33, 130), (2, 100, 1000), (1, 1024, 4096)):
    xyz, ref, ds = case(b, nq, nk, nq)
    os.environ["VDETR_DT_IMPL"] = "3"
    a = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    os.environ["VDETR_DT_IMPL"] = "6"
    c = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    c2 = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    torch.cuda.synchronize()
    print(f"{b}x{nq}x{nk}: max|dt3| {a.abs().max().item():.4e}  max|dt6 - dt3| / max {((c - a).abs().max() / a.abs().max()).item():.3e}  "
          f"finite {bool(torch.isfinite(c).all())}  run-to-run identical {bool(torch.equal(c, c2))}", flush=True)

xyz, ref, ds = case(B, 1024, 4096, 1)
splits = os.environ.get("DT6_SPLITS", "")
if splits:      # developer: time the kernel for several column splits of the two MMAs of a K-step
    for n0 in splits.split(","):
        os.environ["VDETR_DT_IMPL"] = "6"; os.environ["VDETR_DT6_DBG"] = "0"; os.environ["VDETR_DT6_N0"] = n0
        for _ in range(2):
            c = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
        torch.cuda.synchronize()
        import ctypes
        from vdetr_b200 import _C
        _C.lib().vdetr_timing_enable(1)
        for _ in range(5):
            c = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
        torch.cuda.synchronize()
        tot = (ctypes.c_float * 4)(); cnt = (ctypes.c_int * 4)()
        _C.lib().vdetr_timing_read(tot, cnt)
        _C.lib().vdetr_timing_enable(0)
        os.environ["VDETR_DT_IMPL"] = "3"
        a = ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
        print(f"N0 = {n0}: {tot[2] / max(cnt[2], 1):.4f} ms per call; max|dt6 - dt3| / max {((c - a).abs().max() / a.abs().max()).item():.3e}", flush=True)
    sys.exit(0)
for impl, dbg in (("3", "0"), ("6", "0"), ("6", "1"), ("6", "2"), ("6", "3")):
    os.environ["VDETR_DT_IMPL"] = impl
    os.environ["VDETR_DT6_DBG"] = dbg
    for _ in range(2):
        ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    torch.cuda.synchronize()
    import ctypes
    from vdetr_b200 import _C
    _C.lib().vdetr_timing_enable(1)
    for _ in range(5):
        ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    torch.cuda.synchronize()
    tot = (ctypes.c_float * 4)(); cnt = (ctypes.c_int * 4)()
    _C.lib().vdetr_timing_read(tot, cnt)
    print(f"impl {impl} dbg {dbg}: dTables (all kernels of the pass) {tot[2] / max(cnt[2], 1):.4f} ms per call at B={B}", flush=True)
    _C.lib().vdetr_timing_enable(0)

if os.environ.get("VDETR_DT_CLOCKS") == "1":
    os.environ["VDETR_DT_IMPL"] = "6"; os.environ["VDETR_DT6_DBG"] = "0"
    buf = (ctypes.c_ulonglong * 8)()
    _C.lib().vdetr_debug_dt6_clocks(buf)
    for _ in range(4):
        ops.rpe_bias_grad_tables(xyz, ref, None, t0, ds)
    torch.cuda.synchronize()
    _C.check(_C.lib().vdetr_debug_dt6_clocks(buf))
    items = B * 1024 * 64 * 4
    v = [buf[i] / items / (6 if i < 4 else 2) for i in range(6)]
    print("dt6 cycles per producer step: compute %.0f wait-empty %.0f store+fence+arrive %.0f load-issue %.0f | per item in an MMA warp: wait-full %.0f issue+commit %.0f"
          % tuple(v))
    print("per CTA and call: prologue %.0f cycles, producer loop of warp 0 %.0f cycles" % (buf[6] / 148 / 4, buf[7] / 148 / 4))
