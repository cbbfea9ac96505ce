#!/usr/bin/env python
"""bench.py -- scenes/sec, forward+backward, of the V-DETR decoder hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config C3 of BASELINE.md at N=1, C4 at N=8): per GPU a batch of 8 synthetic ScanNet-shaped scenes,
4096 key tokens x 1024 queries x 8 decoder layers, train mode, forward + backward through the whole
TransformerDecoder (fused Vertex-RPE cross attention, fused self attention, FFN, box heads) + AdamW step; for
N > 1 all gradients live in one flat buffer that is summed with ONE NCCL all-reduce per step (parallel.py);
scenes are independent, so the batch is sharded over ranks with no other collective (weak scaling).

One JSON line on rank 0.  `value` = scenes/s with the step's inputs already resident in HBM; `e2e` = the same
step including the pinned-host -> device copy of the inputs and the device -> host read of the loss.

--impl reference: the reference's own decoder (the unmodified /root/reference modules, or their verbatim git-ignored
copy baseline/_ref/ on the GPU box; the pinned port oracle/decoder_torch.py only if neither exists) timed on all host
cores, train mode fp32, on a bounded sample of the same workload: full 8-layer scenes, forward+backward.
"""
import argparse
import json
import os
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NK, NQ, NLAYERS, PER_GPU_BATCH = 4096, 1024, 8, 8
LOSS_KEYS = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits", "angle_residual_normalized")


def synth_scene(B, nK, seed, torch):
    """keys uniform in an 8 x 8 x 3 m room on the 0.04 m voxel lattice, N(0,1) features."""
    g = torch.Generator().manual_seed(1234 + seed)
    xyz = (torch.rand(B, nK, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0]) / 0.04).round() * 0.04
    mins, maxs = xyz.min(1)[0], xyz.max(1)[0]
    sc = maxs - mins
    feat = torch.randn(nK, B, 256, generator=g)
    size = torch.rand(B, nK, 3, generator=g) + 0.3
    return {"xyz": xyz, "feat": feat, "mins": mins, "maxs": maxs, "center_normalized": (xyz - mins[:, None]) / sc[:, None],
            "size_normalized": size / sc[:, None]}


def loss_weights(torch, nq, num_layers, device):
    g = torch.Generator().manual_seed(99)
    shapes = {"sem_cls_logits": 18, "center_normalized": 3, "size_normalized": 3, "angle_logits": 1,
              "angle_residual_normalized": 1}
    out = []
    for li in range(num_layers + 1):
        d = {}
        for k in LOSS_KEYS:
            n = NK if li == 0 else nq
            c = 1 if (li == 0 and k == "sem_cls_logits") else shapes[k]
            d[k] = torch.randn(n, c, generator=g).to(device)
        out.append(d)
    return out


def synthetic_loss(out, weights):
    loss = 0
    for d, w in zip(list(out["aux_outputs"]) + [out["outputs"]], weights):
        for k in LOSS_KEYS:
            loss = loss + (d[k].float() * w[k]).sum()
    return loss


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def build_ours(torch, num_layers=NLAYERS, nq=NQ, dropout=0.1, mlp_dropout=0.3):
    """The reference's training defaults: --dec_dropout 0.1 (main.py:75; attention, residual and FFN dropouts),
    --mlp_dropout 0.3 (main.py:92; box heads)."""
    from vdetr_b200 import vdetr_transformer as vt
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=False)
    first = vt.FFNLayer(d_model=256, dim_feedforward=256, dropout=dropout)
    layer = vt.GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=dropout, pos_for_key=False, args=args)
    torch.manual_seed(0)
    return vt.TransformerDecoder(first, layer, vt.ScanNetBoxConfig(), num_layers=num_layers, decoder_dim=256, mlp_dropout=mlp_dropout,
                                 mlp_norm="bn1d", mlp_act="relu", mlp_sep=True, pos_for_key=False, num_queries=nq,
                                 cls_loss="focalloss_0.25", is_bilable=True, q_content="random", return_intermediate=True,
                                 args=args)


def import_reference():
    """The UNMODIFIED reference modules (models/vdetr_transformer.py, datasets/scannet.py) from /root/reference or its
    verbatim git-ignored copy baseline/_ref/ (baseline/install_ref.py), with the import shims of SURVEY.md Appendix C.
    Returns (TransformerDecoder, GlobalDecoderLayer, FFNLayer, ScannetDatasetConfig, dir) or None."""
    ref = next((p for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))
                if os.path.isdir(os.path.join(p, "models"))), None)
    if ref is None:
        return None
    try:
        sys.path.insert(0, ref)

        def stub(name, **kw):
            m = types.ModuleType(name)
            m.__dict__.update(kw)
            sys.modules[name] = m
            return m
        stub("mmcv"); stub("mmcv.ops", points_in_boxes_all=None); stub("mmcv.ops.furthest_point_sample")
        stub("plyfile", PlyData=None, PlyElement=None); stub("trimesh")
        stub("models").__path__ = [os.path.join(ref, "models")]
        stub("utils").__path__ = [os.path.join(ref, "utils")]
        from datasets.scannet import ScannetDatasetConfig
        from models.vdetr_transformer import TransformerDecoder, GlobalDecoderLayer, FFNLayer
        return TransformerDecoder, GlobalDecoderLayer, FFNLayer, ScannetDatasetConfig, ref
    except Exception as e:  # pragma: no cover
        sys.stderr.write(f"bench.py: reference import failed ({e!r}); using the oracle port\n")
        return None


def build_cpu_reference(torch, num_layers, dropout, mlp_dropout):
    """The reference decoder on the CPU: the unmodified reference when it is importable (kind 'reference'), else the
    pinned port oracle/decoder_torch.py (kind 'port')."""
    got = import_reference()
    torch.manual_seed(0)
    if got is not None:
        TransformerDecoder, GlobalDecoderLayer, FFNLayer, Cfg, ref = got
        args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128, share_selfattn=False)
        first = FFNLayer(d_model=256, dim_feedforward=256, dropout=dropout)
        layer = GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=dropout, pos_for_key=False, args=args)
        dec = TransformerDecoder(first, layer, Cfg(), num_layers=num_layers, decoder_dim=256, mlp_dropout=mlp_dropout, mlp_norm="bn1d",
                                 mlp_act="relu", mlp_sep=True, pos_for_key=False, num_queries=NQ, cls_loss="focalloss_0.25",
                                 is_bilable=True, q_content="random", return_intermediate=True, args=args).train()

        def run(sc, feat):
            out, _ = dec(None, feat, sc["xyz"], sc["xyz"], [sc["mins"], sc["maxs"]], query_pos=None,
                         enc_box_predictions={"center_normalized": sc["center_normalized"], "size_normalized": sc["size_normalized"]},
                         enc_box_features=feat)
            return out
        return dec, run, "reference", ref
    from oracle import decoder_torch as odt
    dec = odt.OracleDecoder(num_layers=num_layers, num_queries=NQ, dropout=dropout, mlp_dropout=mlp_dropout).train()
    for m in dec.modules():
        if isinstance(m, odt.OracleVertexRPECrossAttention):
            m.use_grid_sample = True

    def run(sc, feat):
        out, _ = dec(feat, sc["xyz"], [sc["mins"], sc["maxs"]], sc["center_normalized"], sc["size_normalized"])
        return out
    return dec, run, "port", "oracle/decoder_torch.py"


def cpu_reference_scene_seconds(torch, threads, num_layers, reps, dropout, mlp_dropout, budget_s=None):
    """Forward + backward of ONE scene (4096 keys x 1024 queries, `num_layers` decoder layers + the proposal stage) of the
    reference decoder on the host cores, train mode, fp32: the same synthetic scene, loss and dropout settings as the GPU
    arm.  Returns (list of seconds per repetition, kind, source)."""
    import contextlib
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(sys.stderr):          # the reference prints while it builds; stdout carries the JSON line only
        dec, run, kind, src = build_cpu_reference(torch, num_layers, dropout, mlp_dropout)
    sc = synth_scene(1, NK, 0, torch)
    weights = loss_weights(torch, NQ, num_layers, "cpu")
    feat = sc["feat"].clone().requires_grad_(True)
    times = []
    t_start = time.time()
    for _ in range(reps):
        t0 = time.time()
        with contextlib.redirect_stdout(sys.stderr):
            out = run(sc, feat)
        loss = synthetic_loss(out, weights)
        dec.zero_grad(set_to_none=True)
        feat.grad = None
        loss.backward()
        times.append(time.time() - t0)
        if budget_s is not None and time.time() - t_start + times[-1] > budget_s:
            break
    return times, kind, src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="scenes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1, help="--dec_dropout of the reference (main.py:75)")
    ap.add_argument("--mlp-dropout", type=float, default=0.3, help="--mlp_dropout of the reference (main.py:92)")
    ap.add_argument("--profile", default="", help="write a torch.profiler kernel table of one step to this file and exit")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of replaying a captured CUDA graph")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange fused with AdamW over peer memory (csrc/peer.cu) or one NCCL all-reduce + AdamW")
    ap.add_argument("--no-sync-bn", action="store_true",
                    help="N > 1: keep BatchNorm statistics per GPU (default: SyncBatchNorm over peer memory, as main.py:512-514)")
    ap.add_argument("--max-seconds", type=int, default=480, help="hard watchdog: abort instead of hanging")
    ap.add_argument("--reference-budget", type=int, default=200, help="--impl reference: stop sampling after this many seconds")
    a = ap.parse_args()
    import signal

    def _abort(signum, frame):
        sys.stderr.write("bench.py: watchdog expired, aborting\n")
        os._exit(3)
    signal.signal(signal.SIGALRM, _abort)
    signal.alarm(a.max_seconds)

    import torch
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": f"C3/C4: {a.batch} scenes per GPU x {NK} keys x {NQ} queries x {NLAYERS} decoder layers, "
                          f"fwd+bwd+AdamW, train mode (BatchNorm batch statistics{" synchronised over the GPUs" if (world > 1 and not a.no_sync_bn and a.exchange == "peer") else ""}, dec_dropout {a.dropout} incl. attention dropout "
                          f"inside the fused kernels, mlp_dropout {a.mlp_dropout}), TF32 Linear/Conv layers",
              "global_batch": a.batch * world, "parallelism": f"dp{world}",
              "l2": "per-step working set (> 6 GB of attention scratch, saved bias and activations) >> 126 MB L2; no explicit flush",
              "notes": "training keeps the fused forward's per-pair bias (537 MB per layer at batch 8) for the backward"}

    if a.impl == "reference":
        # The reference's own CPU implementation of the path, all host cores.  A step of the metric is a.batch scenes; the CPU
        # needs about a minute per scene, so every timed step is a bounded sample of the step -- ONE full scene, all 8
        # decoder layers, forward + backward (scenes are independent: the step costs a.batch times the sample) -- and the
        # number of timed samples is cut so that the whole run ends within a few minutes.
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        threads = cores
        times, kind, src = cpu_reference_scene_seconds(torch, threads, NLAYERS, max(1, a.steps), a.dropout, a.mlp_dropout,
                                                       budget_s=a.reference_budget)
        sec_per_scene = sorted(times)[len(times) // 2]
        val = 1.0 / sec_per_scene
        line = {"impl": "reference", "metric": "scenes/sec fwd+bwd, 4096 keys x 1024 queries x 8 dec layers", "value": val,
                "unit": "scenes/s", "n_gpus": a.gpus, "steps": len(times), "warmup": 0,
                "ms_per_step": 1000.0 * sec_per_scene * a.batch,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": threads, "kind": kind,
                                 "sample": f"{len(times)} x (1 of the step's {a.batch} scenes: 4096 keys x 1024 queries, proposal stage + all "
                                           f"{NLAYERS} decoder layers, forward+backward, train mode, fp32) of {src}: "
                                           + ", ".join(f"{t:.1f}" for t in times) + f" s; ms_per_step = {a.batch} x the median sample "
                                           f"(scenes are independent); {a.steps} steps requested, cut to a {a.reference_budget} s budget; "
                                           f"{cores} host cores, {threads} torch threads"},
                "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (vdetr_b200 has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the nn.Linear / Conv1d layers around the kernels run on the TF32 tensor-core path (the configs name bf16)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    ddp = world > 1
    if ddp:
        import torch.distributed as dist
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # the all-reduce is captured in a CUDA graph
        dist.init_process_group("nccl", device_id=dev)
    import vdetr_b200._C as C

    dec = build_ours(torch, dropout=a.dropout, mlp_dropout=a.mlp_dropout).to(dev).train()
    for p in dec.pointcls_heads.parameters():        # used by ModelVDETR.forward, not by the decoder itself
        p.requires_grad_(False)
    from vdetr_b200 import parallel
    model = dec
    use_graph = not a.no_graph and not a.profile
    # AdamW on flat buffers (parallel.FlatAdamW, csrc/optim.cu: one launch per step); its gradient views are the buffer
    # that is all-reduced: ONE NCCL all-reduce per step, the 1 / world factor folded into the optimizer's gradient scale
    peer, exchange, sync_bn = None, "none (1 GPU)", False
    if ddp and a.exchange == "peer":
        try:
            peer = parallel.PeerGroup(dev)
            exchange = "fused reduce-scatter + AdamW + all-gather over peer memory (csrc/peer.cu), no NCCL call in the step"
        except Exception as e:      # e.g. no P2P access between the GPUs of this box: the NCCL form of the same step
            sys.stderr.write(f"bench.py: peer memory unavailable ({e!r}); using the NCCL all-reduce\n")
            peer = None
    if ddp and peer is None:
        exchange = "one NCCL all-reduce of the flat gradient buffer + flat AdamW"
    opt = parallel.FlatAdamW(dec.named_parameters(), lr=1e-5, weight_decay=0.1, peer=peer)
    opt.world_scale = 1.0 / world
    gsync = opt.grads
    if peer is not None and not a.no_sync_bn:
        peer.enable_sync_batchnorm()
        sync_bn = True
    # (how this arm ran: kept out of `config`, which both arms print identically)
    run_info = {"exchange": exchange,
                "batchnorm": "SyncBatchNorm statistics over peer memory" if sync_bn else "batch statistics per GPU"}
    lo, hi = parallel.shard_range(a.batch * world, rank, world)     # scenes [lo, hi) of the global batch live on this rank
    assert hi - lo == a.batch
    host = synth_scene(a.batch, NK, lo, torch)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    weights = loss_weights(torch, NQ, NLAYERS, dev)
    resident = {k: v.to(dev) for k, v in host.items()}

    def fwd_bwd(inp):
        out, _ = model(None, inp["feat"], inp["xyz"], inp["xyz"], [inp["mins"], inp["maxs"]], query_pos=None,
                       enc_box_predictions={"center_normalized": inp["center_normalized"],
                                            "size_normalized": inp["size_normalized"]}, enc_box_features=inp["feat"])
        loss = synthetic_loss(out, weights)
        gsync.zero_()          # .grad = None: backward hands each parameter its gradient without an add_ kernel
        loss.backward()
        gsync.gather_()        # one multi-tensor copy into the flat buffer that is all-reduced / stepped
        return loss

    def step(inp, fetch_loss):
        loss = fwd_bwd(inp)
        if peer is None:
            gsync.sync_(average=False)   # N > 1: the single NCCL all-reduce (SUM) of the flat gradient buffer
        opt.step()                       # peer form: the exchange is inside the optimizer kernels
        return loss.item() if fetch_loss else loss

    def e2e_step():
        inp = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        return step(inp, True)

    def timed(fn, n):
        if ddp:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
        if ddp:
            dist.barrier()
        return ms

    for _ in range(max(a.warmup, 3)):
        step(resident, False)
    torch.cuda.synchronize()
    # ---- eager pass with per-kernel CUDA events (roofline numbers); the headline loop below replays a CUDA graph
    C.lib().vdetr_timing_enable(1)
    C.lib().vdetr_launch_count(1)
    for _ in range(2):
        step(resident, False)
    own_launches_per_step = int(C.lib().vdetr_launch_count(1)) // 2       # kernels of libvdetr_b200 per step
    import ctypes
    tot = (ctypes.c_float * 4)()
    cnt = (ctypes.c_int * 4)()
    C.lib().vdetr_timing_read(tot, cnt)
    C.lib().vdetr_timing_enable(0)
    timed_eager_steps = 2
    graph = None
    if use_graph:
        # whole step (forward, loss, backward, AdamW) captured once; inputs live in static device buffers
        static_in = {k: v.clone() for k, v in resident.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step(static_in, False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # NCCL form: two graphs with the (eager, one launch) all-reduce between them: forward+backward | optimizer.
        # Peer form: the cross-GPU barriers and P2P kernels are ordinary graph nodes, the whole step is two replays back to back.
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = fwd_bwd(static_in)
        graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_opt, pool=graph.pool()):
            opt.step()
        torch.cuda.synchronize()

        def graph_step():
            graph.replay()
            if peer is None:
                gsync.sync_(average=False)
            graph_opt.replay()
            return static_loss

        def e2e_graph_step():
            for k_, v_ in pinned.items():
                static_in[k_].copy_(v_, non_blocking=True)
            graph_step()
            return static_loss.item()
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(resident, False)
            torch.cuda.synchronize()
        os.makedirs(os.path.dirname(os.path.abspath(a.profile)), exist_ok=True)
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=90, max_name_column_width=110))
        return
    sampler = ClockSampler(local)
    sampler.start()
    if graph is not None:
        for _ in range(2):
            graph_step()
        ms_total = timed(graph_step, a.steps)
        for _ in range(2):
            e2e_graph_step()
        ms_e2e = timed(e2e_graph_step, a.steps)
    else:
        ms_total = timed(lambda: step(resident, False), a.steps)
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, a.steps)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if peer is not None:
        peer.check()                 # raises if a cross-GPU barrier of this rank ever timed out: the numbers would be invalid
    scenes = a.batch * world * a.steps
    value = scenes / (ms_total / 1000.0)
    e2e = scenes / (ms_e2e / 1000.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    # fused forward kernel: algorithmic FLOPs = 4*H*nQ*nK*hd per layer-scene (QK^T + PV), BASELINE.md section 3
    flops_fwd = 4.0 * 4 * NQ * NK * 64 * a.batch
    run_info["launch"] = ("CUDA graph replay (fwd+bwd graph, optimizer graph" +
                          (", eager NCCL all-reduce between them)" if (ddp and peer is None) else ")")) if graph is not None \
        else "eager launches"
    fwd_ms = tot[0] / max(cnt[0], 1)
    dt_ms = tot[2] / max(cnt[2], 1)
    achieved = flops_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0
    evals = 8.0 * NQ * NK * a.batch                            # vertex evaluations per launch (forward bias / dTables adjoint)
    pairs = float(NQ) * NK * a.batch
    # ncu --set full captures of the same kernels at this batch (profiles/r2_ncu_traffic.json): DRAM bytes and executed
    # warp-instructions per launch
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
    except Exception:
        pass
    full = a.batch == PER_GPU_BATCH
    peak_bw = peaks.get("hbm_gbs", 6450.0)
    # dTables (adjoint of the bias gather) is the dominant kernel of the step: dt6, a dense tcgen05 contraction.
    # Algorithmic work per pair: 8 vertices x 8 corners x 4 heads = 256 multiply-adds = 512 FLOP; mandatory HBM traffic: the
    # scaled fp16 dS of the 4 heads (8 B per pair) + key xyz + query geometry + one 128 KB table copy per CTA.
    # What the kernel EXECUTES on the tensor pipe is the dense form of that contraction: per 64 pairs 4 K-steps of
    # 128 x (256 + 144) x 16, i.e. 102 400 FLOP per pair of which 512 touch non-zero weights.
    dt_flops = pairs * 512.0
    dt_dense_flops = pairs * (4 * 128 * 400 * 16 * 2) / 64.0
    dt_bytes = pairs * 8 + a.batch * (NK * 16 + NQ * 144) + 148 * 4 * 100 * 80 * 4
    dt_tf = dt_flops / (dt_ms * 1e-3) / 1e12 if dt_ms > 0 else 0.0
    dt_dense_tf = dt_dense_flops / (dt_ms * 1e-3) / 1e12 if dt_ms > 0 else 0.0
    dt_gbs = dt_bytes / (dt_ms * 1e-3) / 1e9 if dt_ms > 0 else 0.0
    dt_inst = ncu.get("dtables", {}).get("warp_inst")
    line = {"metric": "scenes/sec fwd+bwd, 4096 keys x 1024 queries x 8 dec layers", "value": value, "unit": "scenes/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_total / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 tensor-core operands (QK^T from hi/lo fp16 splits of fp32 q and k; P, V, dO, dS as (scaled) fp16), "
                     "fp32 bias / softmax / accumulators; TF32 for the nn.Linear GEMMs",
            "data": "synthetic", "config": config, "run": run_info,
            "e2e": {"value": e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": own_launches_per_step * a.steps,
            "clocks": sampler.summary(),
            "roofline": {"kernel": "dt6::rpe_dtables_umma_kernel + reduction (adjoint of the Vertex-RPE bias gather as a dense tcgen05 "
                                   f"contraction: the dominant kernel, {tot[2] / timed_eager_steps:.1f} ms of the step)",
                         "bound": "tensor", "achieved": dt_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": dt_tf / peak_tf if peak_tf else None,
                         "traffic": ncu.get("dtables", {}).get("dram_bytes") if full else None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                         "launch_ms": dt_ms, "algorithmic_flops": dt_flops, "algorithmic_bytes": dt_bytes,
                         "note": "achieved counts the 512 useful FLOP per (query, key) pair only; the kernel reaches them by executing the "
                                 "dense 128 x 400 MMAs of the factored weights (96 % zeros) because that costs 13x fewer issue slots than "
                                 "sorting + register accumulation (dt3: 3.7 ms) -- see executed_dense_mma for how busy the tensor pipe is",
                         "executed_dense_mma": {"tflops": dt_dense_tf, "frac_of_peak": dt_dense_tf / peak_tf if peak_tf else None,
                                                "frac_of_nominal_dense_fp16_peak": dt_dense_tf / 2250.0,
                                                "flops_per_pair": 102400, "mma_floor_ms": dt_dense_flops / (2 * 4096.0 * 148 * 1.965e9) * 1e3,
                                                "tensor_pipe_active_pct_ncu": ncu.get("dtables", {}).get("tensor_pipe_pct") if full else None},
                         "hbm_view": {"achieved_gbs": dt_gbs, "peak_gbs": peak_bw, "frac": dt_gbs / peak_bw if peak_bw else None},
                         "warp_inst_per_pair": dt_inst / pairs if (dt_inst and full) else None},
            "roofline_attention": {"kernel": "rpe_xattn_fwd_kernel<bias,MQA> (fused Vertex-RPE cross attention, forward: the kernel "
                                             "north_star names)",
                                   "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                                   "frac": achieved / peak_tf if peak_tf else None,
                                   # the write of the per-pair bias kept for the backward (16 B x pairs) dominates; algorithmic minimum 19 MB
                                   "traffic": ncu.get("xattn_fwd", {}).get("dram_bytes") if full else None,
                                   "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                                   "launch_ms": fwd_ms,
                                   "secondary_bound": {"what": "instruction issue of the bias gather (8 vertices x 4 (z,y) corners x one "
                                                               "LDS.128 + fp16->fp32 conversions + FFMA per pair), DESIGN.md 4.1",
                                                       "gevals_per_s": evals / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None}},
            "kernel_ms_per_step": {"xattn_fwd": tot[0] / timed_eager_steps, "xattn_bwd_pass1": tot[1] / timed_eager_steps,
                                   "dtables": tot[2] / timed_eager_steps, "xattn_bwd_pass2_dkdv": tot[3] / timed_eager_steps,
                                   "launches_per_step": [int(c) // timed_eager_steps for c in cnt],
                                   "how": "CUDA events around each launch (vdetr_timing_*), eager pass of 2 steps"}}
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            threads = cores
            try:
                # bounded sample (10-30 s of CPU work): one scene, proposal stage + 2 of the 8 decoder layers, forward+backward;
                # a second run with 1 layer separates the per-layer cost from the proposal stage
                t2, kind, src = cpu_reference_scene_seconds(torch, threads, 2, 1, a.dropout, a.mlp_dropout)
                t1, _, _ = cpu_reference_scene_seconds(torch, threads, 1, 1, a.dropout, a.mlp_dropout)
                per_layer = max(t2[0] - t1[0], 1e-6)
                sec_per_scene = t1[0] + (NLAYERS - 1) * per_layer
                line["cpu_baseline"] = {"value": 1.0 / sec_per_scene, "unit": "scenes/s", "cores": threads, "kind": kind,
                                        "sample": f"one scene fwd+bwd of {src} with 2 decoder layers ({t2[0]:.1f} s) and with 1 ({t1[0]:.1f} s), "
                                                  f"extended linearly to 8 layers; {cores} host cores, {threads} threads; the reference "
                                                  "arm (--impl reference) times full 8-layer scenes"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(line))
    if ddp:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
