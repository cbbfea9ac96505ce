"""Developer diagnostic: per-parameter gradient error of the harsh train golden case (see test_decoder_gpu.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import pytest
import test_decoder_gpu as T
mp = pytest.MonkeyPatch()
gold = dict(np.load(os.path.join(T.G, "decoder_train.npz")))
dec, feat, loss = T._train_case(mp, 0, 0)
print("loss rel", abs(loss.item() - float(gold["loss"])) / abs(float(gold["loss"])))
g = feat.grad.cpu().numpy()
print("dfeat", np.abs(g - gold["dfeat"]).max() / np.abs(gold["dfeat"]).max())
rows = []
for n, p in dec.named_parameters():
    key = "grad." + n
    if key in gold and not n.endswith("k.bias"):
        got = p.grad.cpu().numpy()
        got = got[::16] if got.ndim == 2 and got.shape[0] > 64 else got
        want = gold[key]
        rows.append((np.abs(got - want).max() / (np.abs(want).max() + 1e-12), n))
rows.sort(reverse=True)
print("worst:", [(f"{r:.3f}", n) for r, n in rows[:6]], "median", np.median([r for r, _ in rows]))
