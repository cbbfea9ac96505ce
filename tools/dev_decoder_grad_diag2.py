import os, sys, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tests", "golden"))
import recipe
from oracle import decoder_torch as odt
from test_decoder_gpu import build_product_decoder, _load, _run_product
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
gold = dict(np.load(os.path.join(R, "tests/golden/decoder_train.npz")))
for fi, bi in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
    os.environ["VDETR_B200_IMPL"] = fi; os.environ["VDETR_B200_IMPL_BWD"] = bi
    dec = build_product_decoder(2, 32, dropout=0.0, mlp_dropout=0.0); _load(dec, 41); dec = dec.cuda().train()
    out, feat = _run_product(dec, recipe.decoder_case(42, 2, 96), True)
    loss = odt.synthetic_loss(out); loss.backward()
    g = feat.grad.cpu().numpy(); w = gold["dfeat"]
    print("fwd impl", fi, "bwd impl", bi, "loss", round(loss.item(), 4), "dfeat max-rel", np.abs(g - w).max() / np.abs(w).max(), "rms-rel", np.sqrt(((g - w) ** 2).mean()) / np.sqrt((w ** 2).mean()))
