"""Developer timing of batched FPS (N = 50 000 lattice points, M = 4096) at B = 1 / 8 and of the ragged variant."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vdetr_b200.pointnet2_utils as pu
g = torch.Generator().manual_seed(7)
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
for B in (1, 4, 8):
    pts = ((torch.rand(B, 50000, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04).cuda().contiguous()
    print("fps B", B, "ms", round(t(lambda: pu._ext.furthest_point_sampling(pts, 4096)), 3))
    for plan in ("8,16", "16,8"):
        os.environ["VDETR_FPS_PLAN"] = plan
        try:
            print("   plan", plan, "ms", round(t(lambda: pu._ext.furthest_point_sampling(pts, 4096)), 3))
        except Exception as e:
            print("   plan", plan, "failed", repr(e)[:80])
        del os.environ["VDETR_FPS_PLAN"]
counts = [50000, 43000, 38000, 50000, 47000, 31000, 50000, 45000]
pts = ((torch.rand(sum(counts), 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04).cuda().contiguous()
off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32).cuda()
print("ragged B 8 ms", round(t(lambda: pu._ext.furthest_point_sampling_ragged(pts, off, 4096, 50000)), 3))
