"""Developer script for ncu: one fused forward (+ optional backward) at the C3 shape.  Usage under ncu:
   ncu --set full -k regex:rpe_xattn_fwd -c 1 -o gpurun_out/fwd python tools/dev_ncu_ops.py fwd"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdetr_b200 import ops
from vdetr_b200.vdetr_transformer import morton_order
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nQ, nK = 1024, 4096
g = torch.Generator().manual_seed(0)
xyz = ((torch.rand(B, nK, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04).cuda()
perm = morton_order(xyz)
xyz = torch.gather(xyz, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
center = (torch.rand(B, nQ, 3, generator=g) * torch.tensor([8., 8., 3.])).cuda()
size = (torch.rand(B, nQ, 3, generator=g) + 0.3).cuda()
sgn = torch.tensor([[1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1], [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1]], dtype=torch.float32).cuda()
ref = (center[:, :, None, :] + sgn * size[:, :, None, :] / 2).contiguous()
q = (torch.randn(B, nQ, 4, 64, generator=g) * 0.3).cuda().requires_grad_(True)
k = torch.randn(B, nK, 1, 64, generator=g).cuda().requires_grad_(True)
v = torch.randn(B, nK, 1, 64, generator=g).cuda().requires_grad_(True)
tables = (torch.randn(8, 10, 10, 10, 4, generator=g) * 0.5).cuda().requires_grad_(True)
for it in range(2):
    o = ops.rpe_attention(q, k, v, xyz, ref, None, tables)
    if mode != "fwd":
        o.backward(torch.ones_like(o) * 1e-3)
torch.cuda.synchronize()
import ctypes
from vdetr_b200 import _C
_C.lib().vdetr_timing_enable(1)
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for it in range(5):
    o = ops.rpe_attention(q, k, v, xyz, ref, None, tables)
    if mode != "fwd":
        o.backward(torch.ones_like(o) * 1e-3)
b.record(); torch.cuda.synchronize()
print(mode, "B", B, "ms per call", a.elapsed_time(b) / 5)
tot = (ctypes.c_float * 4)(); cnt = (ctypes.c_int * 4)()
_C.lib().vdetr_timing_read(tot, cnt)
print("kernel ms per launch:", {n: round(t / max(c, 1), 4) for n, t, c in zip(["fwd", "bwd_pass1", "dtables(all kernels)", "dkdv"], tot, cnt)})
if os.environ.get("VDETR_DT_CLOCKS") == "1":
    buf = (ctypes.c_ulonglong * 8)()
    _C.check(_C.lib().vdetr_debug_dt_clocks(buf))
    tot = sum(buf) or 1
    print("dT phase cycles (sum over CTAs, 7 calls):", [f"{n}:{100 * v / tot:.1f}%" for n, v in zip(["A", "zero/B0", "B1", "B2", "B3", "B4", "B4 warp mean", "B4 warp max"], buf)], "total", tot)
