"""Generate golden vectors from the UNMODIFIED reference (needs /root/reference; run in the build
container only):   python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Inputs and weights come from tests/golden/recipe.py (numpy RandomState),
so the fixtures hold outputs and gradients only.  The import shims follow SURVEY.md Appendix C.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipe  # noqa: E402

REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)

    def stub(name, **a):
        m = types.ModuleType(name)
        m.__dict__.update(a)
        sys.modules[name] = m
        return m
    stub("mmcv"); stub("mmcv.ops", points_in_boxes_all=None); stub("mmcv.ops.furthest_point_sample")
    stub("plyfile", PlyData=None, PlyElement=None); stub("trimesh")
    stub("models").__path__ = [os.path.join(REF, "models")]
    stub("utils").__path__ = [os.path.join(REF, "utils")]
    from datasets.scannet import ScannetDatasetConfig
    import models.vdetr_transformer as vt
    return ScannetDatasetConfig, vt


def ref_vertices(cfg, center, size, angle, vt):
    corners = cfg.box_parametrization_to_corners(center, size, angle)
    return vt.convert_corners_camera2lidar(corners.clone())


def gen_xattn(vt, cfg, name, seed, B, nQ, nK, rotated, far):
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10",
                                 angle_type="object_coords" if rotated else "", rpe_dim=128)
    mod = vt.GlobalShareCrossAttention(256, 4, args=args).eval()
    params = recipe.xattn_params(seed)
    sd = mod.state_dict()
    for k, v in params.items():
        sd[k] = torch.from_numpy(v)
    mod.load_state_dict(sd)
    case = recipe.xattn_case(seed + 1, B, nQ, nK, rotated, far)
    center, size = torch.from_numpy(case["center"]), torch.from_numpy(case["size"])
    angle = torch.from_numpy(case["angle"]) if rotated else torch.zeros(B, nQ)
    ref = ref_vertices(cfg, center, size, angle, vt)
    query = torch.from_numpy(case["query"]).requires_grad_(True)
    key = torch.from_numpy(case["key"]).requires_grad_(True)
    x, attn = mod(query, key, ref, angle if rotated else None, torch.from_numpy(case["xyz"]))
    x.backward(torch.from_numpy(case["dout"]))
    out = {"x": x.detach().numpy(), "attn": attn.detach().numpy(), "ref_pts": ref.numpy(),
           "dquery": query.grad.numpy(), "dkey": key.grad.numpy()}
    for n, p in mod.named_parameters():
        g = p.grad.numpy()
        out["grad." + n] = g[::8] if g.shape == (256, 256) else g
    tabs = torch.stack([m(mod.relative_coords_table)[0] for m in mod.cpb_mlps]).detach().numpy()
    out["tables"] = tabs
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("grad.")})


def gen_xattn_mask(vt, cfg, name, seed, B, nQ, nK):
    """attn_mask semantics of the reference (models/vdetr_transformer.py:743-749): a bool mask sets the LOGIT to -100,
    a float mask is added.  Same module / inputs as xattn_small; the masks come from recipe.xattn_masks."""
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128)
    mod = vt.GlobalShareCrossAttention(256, 4, args=args).eval()
    sd = mod.state_dict()
    for k, v in recipe.xattn_params(seed).items():
        sd[k] = torch.from_numpy(v)
    mod.load_state_dict(sd)
    case = recipe.xattn_case(seed + 1, B, nQ, nK, False, 0.3)
    ref = ref_vertices(cfg, torch.from_numpy(case["center"]), torch.from_numpy(case["size"]), torch.zeros(B, nQ), vt)
    mb, mf = recipe.xattn_masks(seed + 2, B, nQ, nK)
    out = {}
    with torch.no_grad():
        for tag, m in (("bool", torch.from_numpy(mb)), ("float", torch.from_numpy(mf))):
            x, attn = mod(torch.from_numpy(case["query"]), torch.from_numpy(case["key"]), ref, None,
                          torch.from_numpy(case["xyz"]), attn_mask=m)
            out["x_" + tag], out["attn_" + tag] = x.numpy(), attn.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def build_ref_decoder(vt, cfg, L, nq, dropout, mlp_dropout, share=False):
    args = types.SimpleNamespace(log_scale=512.0, rpe_quant="bilinear_4_10", angle_type="", rpe_dim=128,
                                 share_selfattn=share)
    first = vt.FFNLayer(d_model=256, dim_feedforward=256, dropout=dropout)
    layer = vt.GlobalDecoderLayer(d_model=256, nhead=4, dim_feedforward=256, dropout=dropout,
                                  pos_for_key=False, args=args)
    return vt.TransformerDecoder(first, layer, cfg, num_layers=L, decoder_dim=256, mlp_dropout=mlp_dropout,
                                 mlp_norm="bn1d", mlp_act="relu", mlp_sep=True, pos_for_key=False,
                                 num_queries=nq, cls_loss="focalloss_0.25", is_bilable=True,
                                 q_content="random", return_intermediate=True, args=args)


OUT_KEYS = ("sem_cls_logits", "center_normalized", "size_normalized", "angle_logits",
            "angle_residual_normalized", "center_unnormalized", "size_unnormalized", "box_corners")


def gen_decoder(vt, cfg, name, seed, B, nK, nq, L, train, share=False):
    torch.manual_seed(0)
    dec = build_ref_decoder(vt, cfg, L, nq, 0.0 if train else 0.1, 0.0 if train else 0.3, share)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    vals = recipe.fill_state_dict(shapes, seed)
    sd = dec.state_dict()
    for k, v in vals.items():
        if v is not None:
            sd[k] = torch.from_numpy(v)
    dec.load_state_dict(sd)
    dec.train(train)
    c = recipe.decoder_case(seed + 1, B, nK)
    feat = torch.from_numpy(c["feat"]).requires_grad_(train)
    xyz = torch.from_numpy(c["xyz"])
    dims = [torch.from_numpy(c["mins"]), torch.from_numpy(c["maxs"])]
    encp = {"center_normalized": torch.from_numpy(c["center_normalized"]),
            "size_normalized": torch.from_numpy(c["size_normalized"])}
    with torch.set_grad_enabled(train):
        o, _ = dec(None, feat, xyz, xyz, dims, query_pos=None, enc_box_predictions=encp,
                   enc_box_features=feat)
    out = {"shapes_keys": np.array(sorted(shapes)), "shapes_vals": np.array([str(shapes[k]) for k in sorted(shapes)])}
    for li, d in enumerate(o["aux_outputs"] + [o["outputs"]]):
        for k in OUT_KEYS:
            out[f"l{li}.{k}"] = d[k].detach().numpy()
    if train:
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from oracle.decoder_torch import synthetic_loss
        loss = synthetic_loss(o)
        loss.backward()
        out["loss"] = np.array(loss.item())
        out["dfeat"] = feat.grad.numpy()
        for n, p in dec.named_parameters():
            if p.grad is None:
                continue
            if "cpb_mlps" in n or n.endswith("norm.weight") or n.endswith("k.bias") or n == "query_embed.weight":
                g = p.grad.numpy()
                out["grad." + n] = g[::16] if g.ndim == 2 and g.shape[0] > 64 else g
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "saved", len(out), "arrays")


if __name__ == "__main__":
    torch.set_num_threads(8)
    Cfg, vt = import_reference()
    cfg = Cfg()
    if "--only-mask" in sys.argv:
        gen_xattn_mask(vt, cfg, "xattn_mask", 11, B=2, nQ=24, nK=80)
        sys.exit(0)
    gen_xattn(vt, cfg, "xattn_small", 11, B=2, nQ=24, nK=80, rotated=False, far=0.3)
    gen_xattn_mask(vt, cfg, "xattn_mask", 11, B=2, nQ=24, nK=80)
    gen_xattn(vt, cfg, "xattn_rot", 21, B=1, nQ=16, nK=48, rotated=True, far=0.2)
    gen_decoder(vt, cfg, "decoder_eval", 31, B=2, nK=96, nq=32, L=2, train=False)
    gen_decoder(vt, cfg, "decoder_train", 41, B=2, nK=96, nq=32, L=2, train=True)
    gen_decoder(vt, cfg, "decoder_share_eval", 51, B=1, nK=64, nq=16, L=1, train=False, share=True)
