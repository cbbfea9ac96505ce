"""Developer micro-benchmark: weight-gradient GEMM dW = dY^T X for token-major Linear layers, plain vs split over the
token dimension (cuBLAS picks 64x64 tiles without split-K for these shapes: 16 CTAs on 148 SMs)."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1000


for T, out, inn in [(8192, 256, 256), (32768, 256, 256), (8192, 18, 256), (8192, 3, 256), (8192, 768, 256), (8192, 64, 256), (32768, 64, 256)]:
    dy = torch.randn(T, out, device="cuda")
    x = torch.randn(T, inn, device="cuda")
    plain = t(lambda: dy.t() @ x)
    res = [f"plain {plain:.1f}us"]
    for C in (4, 8, 16, 32, 64):
        if T // C < 128:
            continue
        f = lambda: torch.bmm(dy.view(C, T // C, out).transpose(1, 2), x.view(C, T // C, inn)).sum(0)
        err = ((dy.t() @ x) - f()).abs().max().item()
        res.append(f"C={C}: {t(f):.1f}us")
    print(T, out, inn, " | ".join(res), f"maxdiff {err:.2e}")
