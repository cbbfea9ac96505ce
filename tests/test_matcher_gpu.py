"""GPU Hungarian matcher (csrc/matcher.cu, vdetr_b200.matcher.Matcher) against scipy.optimize.linear_sum_assignment, the
routine the reference calls on the host (criterion.py:205-228)."""
import types

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu


def _check(cost, nactual, inds, mask):
    B, nQ, ngt = cost.shape
    for b in range(B):
        na = int(nactual[b])
        rows = np.nonzero(mask[b] > 0)[0]
        assert len(rows) == na and len(set(inds[b, rows].tolist())) == na          # every gt matched once
        assert (inds[b][mask[b] == 0] == 0).all()
        if na == 0:
            continue
        r, c = linear_sum_assignment(cost[b, :, :na])
        want = cost[b, r, c].astype(np.float64).sum()
        got = cost[b, rows, inds[b, rows]].astype(np.float64).sum()
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), (b, got, want)
        assert np.array_equal(rows, r) and np.array_equal(inds[b, rows], c)        # generic costs: the optimum is unique


@pytest.mark.parametrize("B,nQ,ngt,seed", [(8, 1024, 64, 0), (3, 256, 64, 1), (2, 1500, 200, 2), (1, 64, 64, 3), (2, 4096, 17, 4)])
def test_lsap_matches_scipy(B, nQ, ngt, seed):
    from vdetr_b200.matcher import linear_sum_assignment_batched
    rs = np.random.RandomState(seed)
    cost = (rs.standard_normal((B, nQ, ngt)) * 3 + rs.rand(B, nQ, 1)).astype(np.float32)
    nactual = rs.randint(0, ngt + 1, size=B)
    nactual[0] = ngt
    if B > 1:
        nactual[1] = 0
    inds, mask = linear_sum_assignment_batched(torch.from_numpy(cost).cuda(), torch.from_numpy(nactual).cuda())
    _check(cost, nactual, inds.cpu().numpy(), mask.cpu().numpy())


def test_lsap_structured_costs_with_ties_reach_the_optimum():
    """Integer-valued costs have many exact ties: the assignment may differ from scipy's, the total cost may not."""
    from vdetr_b200.matcher import linear_sum_assignment_batched
    rs = np.random.RandomState(7)
    cost = rs.randint(0, 6, size=(4, 300, 40)).astype(np.float32)
    inds, mask = linear_sum_assignment_batched(torch.from_numpy(cost).cuda())
    inds, mask = inds.cpu().numpy(), mask.cpu().numpy()
    for b in range(4):
        rows = np.nonzero(mask[b] > 0)[0]
        assert len(rows) == 40 and len(set(inds[b, rows].tolist())) == 40
        r, c = linear_sum_assignment(cost[b])
        assert cost[b, rows, inds[b, rows]].sum() == cost[b, r, c].sum()


def test_matcher_module_equals_reference_formulation():
    """Matcher.forward on synthetic outputs / targets against the reference's formulas evaluated with scipy on the host
    (criterion.py:121-228 restated here: one-hot angle residual, focal class cost, per-scene loop)."""
    from vdetr_b200.matcher import Matcher, huber_loss
    B, nQ, ngt, ncls, nbin = 3, 128, 16, 18, 1
    g = torch.Generator().manual_seed(0)
    out = {"sem_cls_prob": torch.randn(B, nQ, ncls, generator=g), "angle_logits": torch.randn(B, nQ, nbin, generator=g),
           "angle_residual_normalized": torch.randn(B, nQ, nbin, generator=g), "objectness_prob": torch.rand(B, nQ, generator=g),
           "center_reg_dist": torch.rand(B, nQ, ngt, generator=g), "size_reg_dist": torch.rand(B, nQ, ngt, generator=g),
           "gious": torch.rand(B, nQ, ngt, generator=g) * 2 - 1}
    tgt = {"gt_box_sem_cls_label": torch.randint(0, ncls, (B, ngt), generator=g), "gt_angle_class_label": torch.zeros(B, ngt, dtype=torch.int64),
           "gt_angle_residual_label": torch.randn(B, ngt, generator=g) * 0.1, "nactual_gt": torch.tensor([16, 5, 0])}
    args = types.SimpleNamespace(matcher_anglecls_cost=0.3, matcher_anglereg_cost=0.2)
    m = Matcher("focalloss_0.25", cost_class=1.0, cost_objectness=0.0, cost_giou=2.0, cost_center=1.0, cost_size=1.0, args=args)
    res = m({k: v.cuda() for k, v in out.items()}, {k: v.cuda() for k, v in tgt.items()})
    # reference formulation on the host
    p = out["sem_cls_prob"].sigmoid()
    neg = 0.75 * p ** 2 * (-(1 - p + 1e-8).log()); pos = 0.25 * (1 - p) ** 2 * (-(p + 1e-8).log())
    lab = tgt["gt_box_sem_cls_label"].unsqueeze(1).expand(B, nQ, ngt)
    class_mat = torch.gather(pos - neg, 2, lab)
    alab = tgt["gt_angle_class_label"].unsqueeze(1).expand(B, nQ, ngt)
    acls = -torch.gather(out["angle_logits"], 2, alab)
    onehot = torch.zeros(B, nQ, ngt, nbin).scatter_(3, alab.unsqueeze(-1), 1)
    res_gt = (out["angle_residual_normalized"].unsqueeze(2).repeat(1, 1, ngt, 1) * onehot).sum(-1)
    areg = huber_loss(res_gt - (tgt["gt_angle_residual_label"] / (np.pi / nbin)).unsqueeze(1))
    final = (class_mat + 0.0 * (-out["objectness_prob"].unsqueeze(-1)) + out["center_reg_dist"] + 2.0 * (-out["gious"])
             + out["size_reg_dist"] + 0.3 * acls + 0.2 * areg).numpy()
    inds, mask = res["per_prop_gt_inds"].cpu().numpy(), res["proposal_matched_mask"].cpu().numpy()
    for b in range(B):
        na = int(tgt["nactual_gt"][b])
        want_inds, want_mask = np.zeros(nQ, np.int64), np.zeros(nQ, np.float32)
        if na:
            r, c = linear_sum_assignment(final[b, :, :na])
            want_inds[r] = c; want_mask[r] = 1
        assert np.array_equal(inds[b], want_inds) and np.array_equal(mask[b], want_mask)
    assign = res["assignments"]
    assert len(assign) == B and assign[2] == [] and assign[1][0].numel() == 5
