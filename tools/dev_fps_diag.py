"""Developer diagnostic (not a test): where do FPS results diverge?  Run on the GPU box."""
import os, sys, importlib.util
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pn2_util as U
import vdetr_b200.pointnet2_utils as pu
spec = importlib.util.spec_from_file_location("pn2_ref_ext", "oracle/_ref/pn2_ref_ext.so"); ext = importlib.util.module_from_spec(spec); spec.loader.exec_module(ext)
def first_diff(a, b):
    d = np.nonzero(a != b)
    return None if len(d[0]) == 0 else (int(d[0][0]), int(d[1][0]), int(a[d[0][0], d[1][0]]), int(b[d[0][0], d[1][0]]))
for (n, m) in [(3000, 200), (20000, 512), (50000, 512), (50000, 4096)]:
    pts = U.lattice_cloud(7, 1, n)
    t = torch.from_numpy(pts).cuda()
    ref = ext.furthest_point_sampling(t, m).cpu().numpy()
    orc = U.ref_fps(pts, m)
    print(f"N={n} M={m}: ref_ext vs C-oracle first diff:", first_diff(ref, orc), flush=True)
    for plan in ["", "1,24", "2,24", "4,24", "8,8", "8,16", "16,4", "16,8"]:
        if plan:
            os.environ["VDETR_FPS_PLAN"] = plan
        else:
            os.environ.pop("VDETR_FPS_PLAN", None)
        try:
            got = pu.furthest_point_sample(t, m).cpu().numpy()
            print(f"   plan '{plan}': ours vs ref_ext:", first_diff(got, ref), " ours vs oracle:", first_diff(got, orc), flush=True)
        except Exception as e:
            print(f"   plan '{plan}': error {e}", flush=True)
