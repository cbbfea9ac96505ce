import os, sys, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import recipe
from oracle import decoder_torch as odt
from test_decoder_gpu import build_product_decoder, _load, _run_product
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
gold = dict(np.load(os.path.join(R, "tests/golden/decoder_train.npz")))
for impl in ("0", "1"):
    os.environ["VDETR_B200_IMPL"] = impl
    dec = build_product_decoder(2, 32, dropout=0.0, mlp_dropout=0.0); _load(dec, 41); dec = dec.cuda().train()
    out, feat = _run_product(dec, recipe.decoder_case(42, 2, 96), True)
    loss = odt.synthetic_loss(out); loss.backward()
    print("impl", impl, "loss", loss.item(), "gold", float(gold["loss"]))
    g = feat.grad.cpu().numpy(); w = gold["dfeat"]
    print("  dfeat rel err", np.abs(g - w).max() / np.abs(w).max(), "rms rel", np.sqrt(((g - w) ** 2).mean()) / np.sqrt((w ** 2).mean()))
    for li in range(3):
        for k in ("sem_cls_logits", "center_normalized"):
            d = (out["aux_outputs"] + [out["outputs"]])[li][k].detach().cpu().numpy(); ww = gold[f"l{li}.{k}"]
            print(f"  fwd l{li}.{k}", np.abs(d - ww).max() / np.abs(ww).max())
    rows = []
    for n, p in dec.named_parameters():
        key = "grad." + n
        if key in gold:
            got = p.grad.cpu().numpy(); got = got[::16] if got.ndim == 2 and got.shape[0] > 64 else got
            rows.append((np.abs(got - gold[key]).max() / (np.abs(gold[key]).max() + 1e-12), n))
    for r in sorted(rows, reverse=True)[:12]: print("   ", r)
