"""Oracle: LinAssignmentThingsLoss restated on the CPU.  TEST INFRASTRUCTURE ONLY.

Follows loss/lin_assignment_things.py:23-89 (virtual ground truth by linear assignment :23-56, per-image loss :58-89) and
utils/outlier_rejection.py:8-52 (id-range rejection), :57-73 (label centres); scipy.optimize.linear_sum_assignment is the same
solver the reference calls (:46).  Returns the virtual labels too (-1 where a ray is not trained) so that tests can compare the
assignment itself, not only the loss."""
import numpy as np
import scipy.optimize
import torch
import torch.nn.functional as F


def add_position_id_range_cost(cost_matrix, centers_x, frame_min_length=0.3, max_num_inst_at_x=30, id_margin_at_frame_length=30):
    num_ids = cost_matrix.shape[1]
    m = (max_num_inst_at_x + id_margin_at_frame_length) / frame_min_length
    x_limit = (num_ids - id_margin_at_frame_length) / m
    x = (-centers_x + 1) / 2
    lo = torch.clamp(m * (x % x_limit), 0, num_ids - 1).type(torch.long)
    hi = torch.clamp(lo + id_margin_at_frame_length, 0, num_ids - 1)
    ar = torch.arange(num_ids)[None, :]
    ok = torch.logical_and(lo[:, None] <= ar, ar <= hi[:, None])
    cost_matrix[~ok.numpy()] = 10000
    return cost_matrix


def virtual_labels(p_valid, gt_valid, points_valid=None):
    things_mask = gt_valid > 0
    things_gt = gt_valid[things_mask]
    things_prob = p_valid[things_mask][..., 1:]
    labels = sorted(torch.unique(things_gt).tolist())[:things_prob.shape[-1]]
    cost = np.zeros([len(labels), things_prob.shape[-1]])
    for lidx, label in enumerate(labels):
        cost[lidx, :] = -(things_prob[things_gt == label, :].sum(dim=0) / ((things_gt == label).sum() + 1e-4)).numpy()
    if points_valid is not None and len(labels):
        pts = points_valid[things_mask]
        cx = torch.stack([pts[things_gt == label, 0].mean() for label in labels])
        cost = add_position_id_range_cost(cost, cx)
    rows, cols = scipy.optimize.linear_sum_assignment(np.nan_to_num(cost))
    things_labels = torch.zeros_like(things_gt)
    for aidx, lidx in enumerate(rows):
        things_labels[things_gt == labels[lidx]] = int(cols[aidx])
    new_labels = torch.zeros_like(gt_valid)
    new_labels[things_mask] = things_labels + 1
    return new_labels


def lin_assignment_things_loss(p, gt, stuff_mask, points_3d=None):
    """p [B,R,C] (may require grad), gt int64 [B,R], stuff_mask bool [B,R] -> (loss [B,R], virtual labels int64 [B,R])."""
    loss = torch.zeros_like(p[..., 0])
    virt_all = torch.full_like(gt, -1)
    for i in range(p.shape[0]):
        valid = torch.logical_or(stuff_mask[i], gt[i] > 0)
        pv, gv = p[i][valid], gt[i][valid]
        with torch.no_grad():
            virt = virtual_labels(pv.detach(), gv, points_3d[i][valid] if points_3d is not None else None)
        virt_all[i][valid] = virt
        if torch.any(virt != pv.argmax(dim=-1)):
            li = torch.zeros_like(loss[i])
            li[valid] = F.nll_loss(torch.log(pv + 1e-27), virt, reduction='none')
            loss = loss + torch.nn.functional.one_hot(torch.tensor(i), p.shape[0]).to(loss.dtype)[:, None] * li[None, :]
    return loss, virt_all
