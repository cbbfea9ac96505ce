"""Developer check: run-to-run bit determinism of the fused backward in both bias modes (saved / recompute)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import test_xattn_gpu as T
N = int(os.environ.get("DET_RUNS", "40"))
for (seed, B, nQ, nK) in [(51, 2, 70, 333), (53, 1, 256, 1024)][: int(os.environ.get("DET_CASES", "2"))]:
    I = T._core_inputs(seed, B, nQ, nK, 1, False)
    for mode in os.environ.get("DET_MODES", "1,0").split(","):
        os.environ["VDETR_B200_SAVE_BIAS"] = mode
        base = T._run(I, impl=0)
        bad = {k: 0 for k in ("o", "dq", "dk", "dv")}
        worst = {k: 0.0 for k in bad}
        for it in range(N):
            r = T._run(I, impl=0)
            for k in bad:
                if not np.array_equal(r[k], base[k]):
                    bad[k] += 1
                    worst[k] = max(worst[k], float(np.abs(r[k] - base[k]).max() / np.abs(base[k]).max()))
        print(f"case {seed} save_bias={mode}: mismatching runs of {N}:", bad, "worst rel diff", worst, flush=True)

if os.environ.get("DET_MHA", "1") == "1":
    for (seed, B, nQ, nK) in [(61, 2, 70, 333), (62, 1, 130, 260)]:
        I = T._core_inputs(seed, B, nQ, nK, 4, False)
        base = T._run(I, impl=0, has_bias=False)
        bad = {k: 0 for k in ("o", "dq", "dk", "dv")}
        for it in range(N):
            r = T._run(I, impl=0, has_bias=False)
            for k in bad:
                bad[k] += int(not np.array_equal(r[k], base[k]))
        print(f"MHA no-bias case {seed}: mismatching runs of {N}:", bad, flush=True)
