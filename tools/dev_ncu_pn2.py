"""Developer script for ncu: the pointnet2 seed ops at the SURVEY 8(d) sizes (FPS N=50 000 -> M=4096; ball query
N=20 000, M=2048, nsample 64; gather / group).  ncu --set full -k regex:"fps|ball|gather|group" python tools/dev_ncu_pn2.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vdetr_b200.pointnet2_utils as pu

g = torch.Generator().manual_seed(0)
for B in (1, 8):
    xyz = ((torch.rand(B, 50000, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04 + 0.5).cuda().contiguous()
    for _ in range(2):
        idx = pu.furthest_point_sample(xyz, 4096)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx = pu.furthest_point_sample(xyz, 4096); b.record(); torch.cuda.synchronize()
    print(f"fps B={B} N=50000 M=4096: {a.elapsed_time(b):.3f} ms")
xyz = ((torch.rand(1, 20000, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04 + 0.5).cuda().contiguous()
ctr = xyz[:, :2048].contiguous()
feat = torch.randn(1, 128, 20000, generator=g).cuda()
for _ in range(2):
    nb = pu.ball_query(0.2, 64, xyz, ctr)
    grp = pu.grouping_operation(feat, nb)
    gat = pu.gather_operation(feat, nb[:, :, 0].contiguous())
torch.cuda.synchronize()
for name, fn in (("ball_query r=0.2 ns=64", lambda: pu.ball_query(0.2, 64, xyz, ctr)), ("group C=128", lambda: pu.grouping_operation(feat, nb)),
                 ("gather C=128", lambda: pu.gather_operation(feat, nb[:, :, 0].contiguous()))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b):.3f} ms")
