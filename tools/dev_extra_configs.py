"""Developer measurement of the BASELINE configs that are not the headline bench line (C2 forward-only, C5 dense
stress); writes JSON lines.  Usage: python tools/dev_extra_configs.py > gpurun_out/extra.json"""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
dev = torch.device("cuda", 0)


def run(tag, B, nK, nQ, L, iters=10):
    dec = bench.build_ours(torch, num_layers=L, nq=nQ).to(dev).eval()
    sc = {k: v.to(dev) for k, v in bench.synth_scene(B, nK, 0, torch).items()}

    def fwd():
        with torch.no_grad():
            out, _ = dec(None, sc["feat"], sc["xyz"], sc["xyz"], [sc["mins"], sc["maxs"]], query_pos=None,
                         enc_box_predictions={"center_normalized": sc["center_normalized"], "size_normalized": sc["size_normalized"]},
                         enc_box_features=sc["feat"])
        return out["outputs"]["sem_cls_logits"]

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters
    eager = timeit(fwd)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fwd()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd()
    graph = timeit(g.replay)
    print(json.dumps({"config": tag, "B": B, "nK": nK, "nQ": nQ, "layers": L, "mode": "forward, eval, no_grad",
                      "ms_eager": eager, "ms_graph": graph, "scenes_per_s_eager": B * 1000 / eager,
                      "scenes_per_s_graph": B * 1000 / graph, "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    del dec, g
    torch.cuda.empty_cache()


run("C2 (ScanNet-shape single scene, forward)", 1, 4096, 1024, 8)
run("C2 at batch 8 (forward)", 8, 4096, 1024, 8)
run("C5 (dense stress, forward)", 1, 16384, 2048, 12, iters=5)
