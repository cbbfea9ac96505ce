"""Drop-in for ``Matcher`` of /root/reference/criterion.py:100-228: the Hungarian matching between
the queries and the ground-truth boxes of every scene.

Same constructor, same ``forward(outputs, targets)`` and the same two tensors the losses consume
(``per_prop_gt_inds`` [B,nQ] int64, ``proposal_matched_mask`` [B,nQ] float32, criterion.py:302-523).  What differs:
the cost matrix never leaves the GPU -- the reference copies it to the host and calls scipy's
``linear_sum_assignment`` per scene (``final_cost.detach().cpu().numpy()``, :205-218), nine times per training step --
the assignment runs in ONE kernel launch for the whole batch (csrc/matcher.cu, the same shortest-augmenting-path
algorithm as scipy, FP64 duals) and nothing synchronises with the host.  ``assignments`` (the per-scene index lists the
reference also returns, unused by its losses) is built lazily on first access, because its shapes need a host read.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _C


def huber_loss(error, delta=1.0):
    """utils/misc.py:25-36."""
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    return 0.5 * quadratic ** 2 + delta * (abs_error - quadratic)


def linear_sum_assignment_batched(cost, nactual_gt=None):
    """cost [B,nQ,ngt] float32 CUDA (ngt <= nQ <= 4096, ngt <= 512), nactual_gt [B] integer tensor or None (all ngt columns).
    Scene b: minimise sum_g cost[b, q_g, g] over distinct queries q_g for g < nactual_gt[b] (scipy.optimize.
    linear_sum_assignment on cost[b, :, :nactual_gt[b]]).  Returns per_prop_gt_inds [B,nQ] int64 (0 where unmatched) and
    proposal_matched_mask [B,nQ] float32."""
    _C.require_cuda("cost", cost, torch.float32)
    if cost.dim() != 3:
        raise RuntimeError("cost must be [B,nQ,ngt]")
    B, nQ, ngt = cost.shape
    na = None
    if nactual_gt is not None:
        na = nactual_gt.to(device=cost.device, dtype=torch.int32).contiguous()
        if na.numel() != B:
            raise RuntimeError("nactual_gt must have one entry per scene")
    inds = torch.empty(B, nQ, dtype=torch.int64, device=cost.device)
    mask = torch.empty(B, nQ, dtype=torch.float32, device=cost.device)
    with torch.cuda.device(cost.device):
        _C.check(_C.lib().vdetr_lsap(_C.ptr(cost), _C.ptr(na), B, nQ, ngt, _C.ptr(inds), _C.ptr(mask), _C.stream_ptr()))
    return inds, mask


class _LazyAssignments(dict):
    """The matcher's result dict; ``["assignments"]`` (list of [query_idx, gt_idx] per scene, criterion.py:211-224) is
    materialised on first access -- it is the only entry whose shapes depend on device data."""

    def __missing__(self, key):
        if key != "assignments":
            raise KeyError(key)
        inds, mask = self["per_prop_gt_inds"], self["proposal_matched_mask"]
        out = []
        for b in range(inds.shape[0]):
            rows = torch.nonzero(mask[b] > 0, as_tuple=False).squeeze(1)
            out.append([rows, inds[b, rows]] if rows.numel() else [])
        self[key] = out
        return out


class Matcher(nn.Module):
    def __init__(self, cls_loss, cost_class, cost_objectness, cost_giou, cost_center, cost_size, args):
        super().__init__()
        self.cls_loss = cls_loss
        self.cost_class = cost_class
        self.cost_objectness = cost_objectness
        self.cost_giou = cost_giou
        self.cost_center = cost_center
        self.cost_size = cost_size
        self.matcher_anglecls_cost = args.matcher_anglecls_cost
        self.matcher_anglereg_cost = args.matcher_anglereg_cost

    @torch.no_grad()
    def cost_matrix(self, outputs, targets):
        """final_cost [B,nQ,ngt] (criterion.py:121-203), every term as the reference computes it."""
        B, nQ = outputs["sem_cls_prob"].shape[:2]
        ngt = targets["gt_box_sem_cls_label"].shape[1]
        labels = targets["gt_box_sem_cls_label"].unsqueeze(1).expand(B, nQ, ngt)
        if self.cls_loss.split("_")[0] == "focalloss":
            p = outputs["sem_cls_prob"].sigmoid()
            alpha, gamma = 0.25, 2.0
            neg = (1 - alpha) * (p ** gamma) * (-(1 - p + 1e-8).log())
            pos = alpha * ((1 - p) ** gamma) * (-(p + 1e-8).log())
            class_mat = torch.gather(pos - neg, 2, labels)
        else:
            class_mat = -torch.gather(outputs["sem_cls_prob"], 2, labels)
        angle_labels = targets["gt_angle_class_label"].unsqueeze(1).expand(B, nQ, ngt)
        angle_class_mat = -torch.gather(outputs["angle_logits"], 2, angle_labels)
        res = outputs["angle_residual_normalized"]
        gt_res_norm = targets["gt_angle_residual_label"] / (np.pi / res.shape[-1])
        # residual of the ground-truth angle class (the reference builds a one-hot [B,nQ,ngt,nbin] tensor for this, :160-172)
        res_for_gt = torch.gather(res, 2, angle_labels)
        angle_reg_mat = huber_loss(res_for_gt - gt_res_norm.unsqueeze(1), delta=1.0)
        objectness_mat = -outputs["objectness_prob"].unsqueeze(-1)
        return (self.cost_class * class_mat + self.cost_objectness * objectness_mat
                + self.cost_center * outputs["center_reg_dist"] + self.cost_giou * (-outputs["gious"])
                + self.cost_size * outputs["size_reg_dist"] + self.matcher_anglecls_cost * angle_class_mat
                + self.matcher_anglereg_cost * angle_reg_mat)

    @torch.no_grad()
    def forward(self, outputs, targets):
        cost = self.cost_matrix(outputs, targets).float().contiguous()
        inds, mask = linear_sum_assignment_batched(cost, targets["nactual_gt"])
        return _LazyAssignments(per_prop_gt_inds=inds, proposal_matched_mask=mask)
