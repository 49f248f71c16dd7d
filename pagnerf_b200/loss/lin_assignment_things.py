"""LinAssignmentThingsLoss on the device: drop-in for loss/lin_assignment_things.py:13-89 (constructed at pc_nerf/trainer.py:69-80,
called at :484-520).  Same constructor (`outlier_rejection, min_distance, max_distance`), same `forward(inst_probabilities
[B,R,C], labels_gt [B,R], stuff_mask [B,R], points_3d=None) -> loss [B,R]`; the label sort, the cost matrix, the linear
assignment, the relabelling, the arg-max check and the NLL run as five kernels of csrc/loss.cu with no host synchronisation
(the reference does a `.cpu()` per label, scipy on the host and two Python loops per image).  Optimal assignments are unique
unless two costs tie exactly, so the virtual labels equal scipy's; the loss is the same float32 expression."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._lib import call, ptr


class _InstAssignLossFn(Function):
    @staticmethod
    def forward(ctx, p, gt, stuff, points, frame_min_length, max_num_inst_at_x, id_margin):
        if not p.is_cuda:
            raise RuntimeError("pagnerf_b200 losses run on CUDA tensors only (no CPU fallback)")
        pf = p.detach().to(torch.float32).contiguous()
        B, R, C = pf.shape
        g = gt.to(torch.int64).contiguous()
        s = stuff.to(torch.uint8).contiguous()
        pts = points.detach().to(torch.float32).contiguous() if points is not None else None
        dev, m = pf.device, C - 1
        i32, f32 = torch.int32, torch.float32
        labels = torch.empty(B, m, dtype=i32, device=dev)
        n_labels = torch.empty(B, dtype=i32, device=dev)
        rank = torch.empty(B, R, dtype=i32, device=dev)
        acc = torch.zeros(B * m * m + 2 * B * m, dtype=f32, device=dev)      # csum | cnt | xsum: one memset
        csum, cnt, xsum = acc[:B * m * m], acc[B * m * m:B * m * m + B * m], acc[B * m * m + B * m:]
        assign = torch.zeros(B, m, dtype=i32, device=dev)
        virt = torch.empty(B, R, dtype=i32, device=dev)
        flag = torch.zeros(B, dtype=i32, device=dev)
        loss = torch.empty(B, R, dtype=f32, device=dev)
        call("pag_inst_assignment_loss_fwd", ptr(pf), ptr(g), ptr(s), ptr(pts), B, R, C, float(frame_min_length), int(max_num_inst_at_x),
             int(id_margin), ptr(labels), ptr(n_labels), ptr(rank), ptr(csum), ptr(cnt), ptr(xsum), ptr(assign), ptr(virt), ptr(flag), ptr(loss))
        ctx.save_for_backward(pf, virt, flag)
        ctx.mark_non_differentiable(virt)
        return loss, virt

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loss, _g_virt):
        pf, virt, flag = ctx.saved_tensors
        B, R, C = pf.shape
        gp = torch.zeros_like(pf)
        call("pag_inst_assignment_loss_bwd", ptr(pf), ptr(virt), ptr(flag), ptr(g_loss.to(torch.float32).contiguous()), B, R, C, ptr(gp))
        return gp, None, None, None, None, None, None


class LinAssignmentThingsLoss(nn.Module):
    def __init__(self, outlier_rejection=False, min_distance=0.2, max_distance=0.5, *args, **kwargs):
        super().__init__()
        self.outlier_rejection = outlier_rejection
        self.min_distance = min_distance
        self.max_distance = max_distance
        # defaults of utils/outlier_rejection.add_position_id_range_cost (:8-12)
        self.frame_min_length, self.max_num_inst_at_x, self.id_margin_at_frame_length = 0.3, 30, 30
        self.last_virtual_labels = None      # i32 [B, R]: virtual label per ray, -1 where the ray is not trained (diagnostics / tests)

    def forward(self, inst_probabilities, labels_gt, stuff_mask, points_3d=None, *args, **kwargs):
        assert self.outlier_rejection and points_3d is not None or not self.outlier_rejection, 'Outlier rejection requires 3d points'
        pts = points_3d if self.outlier_rejection else None
        loss, virt = _InstAssignLossFn.apply(inst_probabilities, labels_gt, stuff_mask, pts, self.frame_min_length,
                                             self.max_num_inst_at_x, self.id_margin_at_frame_length)
        self.last_virtual_labels = virt
        return loss
