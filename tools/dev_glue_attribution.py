"""Developer script: which Python lines of the package launch the non-attention ("glue") kernels of a training step.

One eager step of bench.py's workload under torch.profiler with stacks; CUDA time of every aten / autograd op is
grouped by (op, innermost frame inside v-detr_b200 or bench.py) for the forward, and by (op, autograd node) for the backward.
    python tools/dev_glue_attribution.py [out.txt]
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
dec = bench.build_ours(torch).to(dev).train()
for p in dec.pointcls_heads.parameters():
    p.requires_grad_(False)
from vdetr_b200 import parallel  # noqa: E402
opt = parallel.FlatAdamW(dec.named_parameters(), lr=1e-5, weight_decay=0.1)
host = bench.synth_scene(bench.PER_GPU_BATCH, bench.NK, 0, torch)
inp = {k: v.to(dev) for k, v in host.items()}
weights = bench.loss_weights(torch, bench.NQ, bench.NLAYERS, dev)


def step():
    out, _ = dec(None, inp["feat"], inp["xyz"], inp["xyz"], [inp["mins"], inp["maxs"]], query_pos=None,
                 enc_box_predictions={"center_normalized": inp["center_normalized"], "size_normalized": inp["size_normalized"]},
                 enc_box_features=inp["feat"])
    loss = bench.synthetic_loss(out, weights)
    opt.grads.zero_()
    loss.backward()
    opt.grads.gather_()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()

groups = collections.defaultdict(lambda: [0.0, 0, set()])
for ev in prof.events():
    t = getattr(ev, "self_device_time_total", 0.0)
    if t <= 0 or not ev.name.startswith(("aten::", "Memcpy", "Memset")):
        continue
    where = "?"
    for fr in (ev.stack or []):
        if "v-detr_b200" in fr or "bench.py" in fr or "vdetr_b200" in fr:
            where = fr.split("/")[-1]
            break
    if where == "?":
        # backward ops have no Python stack: name the autograd node they run under
        p = ev.cpu_parent
        while p is not None:
            if "Backward" in p.name or p.name.startswith("autograd::"):
                where = p.name
                break
            p = p.cpu_parent
    g = groups[(ev.name, where)]
    g[0] += t
    g[1] += 1
    g[2].add(str(ev.input_shapes)[:90])
rows = sorted(groups.items(), key=lambda kv: -kv[1][0])
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
tot = sum(v[0] for _, v in rows)
out.write(f"aten-level device time of one step: {tot / 1e3:.2f} ms\n")
for (name, where), (t, n, shapes) in rows[:90]:
    out.write(f"{t / 1e3:8.3f} ms {n:5d}  {name:28s} {where[:70]:70s} {sorted(shapes)[0] if shapes else ''}\n")
