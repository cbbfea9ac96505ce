"""Times the UNMODIFIED reference (copied, git-ignored, under baseline/_ref/) on the GPU box: the reference's own
eager-PyTorch GPU forward / forward+backward of the decoder and of one GlobalShareCrossAttention, its CPU forward,
and the rebuilt reference pointnet2 extension (oracle/_ref).  Reported comparator only -- not product code.
Writes gpurun_out/ref_gpu_baseline.json."""
import glob, importlib.util, json, os, subprocess, sys, time, types

out = {}
os.makedirs("gpurun_out", exist_ok=True)
def dump(): json.dump(out, open("gpurun_out/ref_gpu_baseline.json", "w"), indent=1, default=str)
REF = next((p for p in ["/root/reference", "baseline/_ref"] if os.path.isdir(os.path.join(p, "models"))), None)
out["ref_dir"] = REF
try:
    out["nvidia_smi"] = subprocess.run(["nvidia-smi", "--query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.draw", "--format=csv"],
                                       capture_output=True, text=True).stdout
except Exception as e:
    out["nvidia_smi"] = repr(e)
out["nproc"] = os.cpu_count()
dump()
if REF is None:
    print("NO REFERENCE DIR"); sys.exit(0)
sys.path.insert(0, REF)
def stub(name, **a):
    m = types.ModuleType(name); m.__dict__.update(a); sys.modules[name] = m; return m
stub("mmcv"); stub("mmcv.ops", points_in_boxes_all=None); stub("mmcv.ops.furthest_point_sample")
stub("plyfile", PlyData=None, PlyElement=None); stub("trimesh")
stub("models").__path__ = [os.path.join(REF, "models")]; stub("utils").__path__ = [os.path.join(REF, "utils")]
import torch, warnings
warnings.filterwarnings("ignore")
from datasets.scannet import ScannetDatasetConfig
from models.vdetr_transformer import TransformerDecoder, GlobalDecoderLayer, FFNLayer, GlobalShareCrossAttention
out["torch"] = torch.__version__; out["gpu_name"] = torch.cuda.get_device_name(0)
dev = "cuda"
args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=False)
cfg = ScannetDatasetConfig()
def build(L, nq):
    first = FFNLayer(d_model=256, dim_feedforward=256, dropout=0.1)
    layer = GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=0.1, pos_for_key=False, args=args)
    return TransformerDecoder(first, layer, cfg, num_layers=L, decoder_dim=256, mlp_dropout=0.3, mlp_norm="bn1d", mlp_act="relu",
        mlp_sep=True, pos_for_key=False, num_queries=nq, cls_loss="focalloss_0.25", is_bilable=True, q_content="random",
        return_intermediate=True, args=args)
def scene(B, nK, d):
    g = torch.Generator().manual_seed(1234)
    xyz = (torch.rand(B, nK, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04
    mins, maxs = xyz.min(1)[0], xyz.max(1)[0]; sc = maxs - mins
    feat = torch.randn(nK, B, 256, generator=g); size = torch.rand(B, nK, 3, generator=g) + 0.3
    encp = {"center_normalized": ((xyz - mins[:, None]) / sc[:, None]).to(d), "size_normalized": (size / sc[:, None]).to(d)}
    return xyz.to(d), feat.to(d), [mins.to(d), maxs.to(d)], encp
def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
def run_dec(tag, B, nK, nQ, L, train):
    try:
        torch.manual_seed(0); dec = build(L, nQ).to(dev); dec.train(train)
        xyz, feat, dims, encp = scene(B, nK, dev)
        if train: feat.requires_grad_(True)
        def f():
            with torch.set_grad_enabled(train):
                o, _ = dec(None, feat, xyz, xyz, dims, query_pos=None, enc_box_predictions=encp, enc_box_features=feat)
                if train:
                    loss = 0
                    for d in [o["outputs"]] + o["aux_outputs"]:
                        for k in ["sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual_normalized"]:
                            loss = loss + d[k].float().sum()
                    dec.zero_grad(set_to_none=True); loss.backward()
        torch.cuda.reset_peak_memory_stats(); ms = timeit(f, warm=1 if train else 2, it=3 if train else 5)
        out[tag] = {"ms": ms, "scenes_per_s": B * 1000.0 / ms, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    except Exception as e:
        out[tag] = {"error": repr(e)[:300]}
    print(tag, out[tag], flush=True); dump(); torch.cuda.empty_cache()
run_dec("C1_fwd_fp32_gpu_B1_512x128x1", 1, 512, 128, 1, False)
run_dec("C2_fwd_fp32_gpu_B1_4096x1024x8", 1, 4096, 1024, 8, False)
run_dec("C2_fwd_fp32_gpu_B8_4096x1024x8", 8, 4096, 1024, 8, False)
run_dec("C3_fwdbwd_fp32_gpu_B1_4096x1024x8", 1, 4096, 1024, 8, True)
run_dec("C3_fwdbwd_fp32_gpu_B8_4096x1024x8", 8, 4096, 1024, 8, True)
try:
    torch.manual_seed(0); m = GlobalShareCrossAttention(256, 4, args=args).to(dev).eval()
    q = torch.randn(1024, 1, 256, device=dev); k = torch.randn(4096, 1, 256, device=dev)
    xyz = torch.rand(1, 4096, 3, device=dev) * 8; ref = torch.rand(1, 1024, 8, 3, device=dev) * 8
    with torch.no_grad():
        out["xattn_fwd_fp32_B1_ms"] = timeit(lambda: m(q, k, ref, None, xyz), it=10)
    qg = q.clone().requires_grad_(True); kg = k.clone().requires_grad_(True)
    def fb():
        m.zero_grad(set_to_none=True); x, _ = m(qg, kg, ref, None, xyz); x.sum().backward()
    out["xattn_fwdbwd_fp32_B1_ms"] = timeit(fb, warm=1, it=3)
    print({k_: v for k_, v in out.items() if k_.startswith("xattn")}, flush=True)
except Exception as e:
    out["xattn_error"] = repr(e)[:300]
dump()
try:
    so = "oracle/_ref/pn2_ref_ext.so"
    spec = importlib.util.spec_from_file_location("pn2_ref_ext", so); _ext = importlib.util.module_from_spec(spec); spec.loader.exec_module(_ext)
    g = torch.Generator().manual_seed(7)
    for N in (20000, 50000):
        pts = ((torch.rand(1, N, 3, generator=g) * torch.tensor([8., 8., 3.]) / 0.04).round() * 0.04).to(dev).contiguous()
        out[f"fps_N{N}_M4096_B1_ms"] = timeit(lambda: _ext.furthest_point_sampling(pts, 4096), it=5)
    pts8 = pts.repeat(8, 1, 1).contiguous(); out["fps_N50000_M4096_B8_ms"] = timeit(lambda: _ext.furthest_point_sampling(pts8, 4096), it=3)
    idx = _ext.furthest_point_sampling(pts, 4096); feats = torch.randn(1, 256, 50000, device=dev)
    out["gather_C256_N50000_M4096_ms"] = timeit(lambda: _ext.gather_points(feats, idx), it=10)
    xyz20 = pts[:, :20000].contiguous(); ctr = xyz20[:, :2048].contiguous()
    for ns in (16, 64): out[f"ballquery_N20000_M2048_ns{ns}_ms"] = timeit(lambda: _ext.ball_query(ctr, xyz20, 0.2, ns), it=5)
    bi = _ext.ball_query(ctr, xyz20, 0.2, 64); f2 = torch.randn(1, 128, 20000, device=dev)
    out["group_C128_M2048_ns64_ms"] = timeit(lambda: _ext.group_points(f2, bi), it=10)
    # ours, same inputs
    sys.path.insert(0, os.getcwd())
    import vdetr_b200.pointnet2_utils as pu
    out["ours_fps_N50000_M4096_B1_ms"] = timeit(lambda: pu._ext.furthest_point_sampling(pts, 4096), it=5)
    out["ours_fps_N50000_M4096_B8_ms"] = timeit(lambda: pu._ext.furthest_point_sampling(pts8, 4096), it=3)
    out["ours_ballquery_N20000_M2048_ns64_ms"] = timeit(lambda: pu._ext.ball_query(ctr, xyz20, 0.2, 64), it=5)
    out["ours_gather_ms"] = timeit(lambda: pu._ext.gather_points(feats, idx), it=10)
    out["ours_group_ms"] = timeit(lambda: pu._ext.group_points(f2, bi), it=10)
    out["fps_match_50000"] = bool((pu._ext.furthest_point_sampling(pts, 4096) == idx).all())
    print({k_: v for k_, v in out.items() if k_.startswith(("fps", "gather", "ball", "group", "ours"))}, flush=True)
except Exception as e:
    out["pn2_error"] = repr(e)[:500]
dump()
try:
    torch.set_num_threads(os.cpu_count()); torch.manual_seed(0); dec = build(8, 1024).eval()
    xyz, feat, dims, encp = scene(1, 4096, "cpu")
    ts = []
    with torch.no_grad():
        for _ in range(2):
            t = time.time(); dec(None, feat, xyz, xyz, dims, query_pos=None, enc_box_predictions=encp, enc_box_features=feat); ts.append(time.time() - t)
    out["C2_fwd_fp32_CPU_B1"] = {"s": min(ts), "threads": torch.get_num_threads(), "cores": os.cpu_count()}
    print("CPU", out["C2_fwd_fp32_CPU_B1"], flush=True)
except Exception as e:
    out["C2_fwd_fp32_CPU_B1"] = {"error": repr(e)[:300]}
dump(); print("DONE")
