import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdetr_b200 import ops
from vdetr_b200.vdetr_transformer import MultiheadSelfAttention, _dense_attention
torch.manual_seed(0)
def dense(q, k, v):
    o, _ = _dense_attention(q, k, v, None, torch.nn.Identity())
    return o
for (B, nQ, nK, kvh) in [(2, 32, 32, 4), (1, 128, 128, 4), (1, 40, 200, 4), (2, 32, 32, 1), (1, 128, 128, 1), (1, 64, 64, 4), (1, 32, 256, 4), (1, 200, 40, 4)]:
    q = torch.randn(B, nQ, 4, 64, device="cuda") * 0.3
    k = torch.randn(B, nK, kvh, 64, device="cuda")
    v = torch.randn(B, nK, kvh, 64, device="cuda")
    with torch.no_grad():
        a = ops.rpe_attention(q, k, v, impl=0)
        s = ops.rpe_attention(q, k, v, impl=1)
        d = dense(q, k.expand(-1, -1, 4, -1) if kvh == 1 else k, v.expand(-1, -1, 4, -1) if kvh == 1 else v)
    print((B, nQ, nK, kvh), "tc-dense", (a - d).abs().max().item(), "simt-dense", (s - d).abs().max().item(),
          "per-head tc err", [(a - d)[:, :, h].abs().max().item() for h in range(4)], flush=True)
m = MultiheadSelfAttention(256, 4).cuda().eval()
x = torch.randn(32, 2, 256, device="cuda")
pos = torch.randn(32, 2, 256, device="cuda")
with torch.no_grad():
    qk = x + pos
    y, _ = m(qk, qk, value=x)
    ref = torch.nn.MultiheadAttention(256, 4).cuda().eval()
    ref.load_state_dict(m.state_dict())
    yr, _ = ref(qk, qk, value=x)
print("module vs nn.MultiheadAttention", (y - yr).abs().max().item())
