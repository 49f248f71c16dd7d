"""Oracle: BAPipeline.transform_rays (pc_nerf/ba_pipeline.py:85-92) on the CPU.  TEST INFRASTRUCTURE ONLY.

The reference switches kaolin's `Camera.extrinsics` to the 'matrix_6dof_rotation' backend (:44) and registers its parameter
tensor [n_cameras, 9] as `camera_extrinsics` (:49-51); `inv_transform_rays` (:88) maps camera-space base rays to world space and
the reference renormalises the directions (:89).  kaolin is an unpinned, un-vendored dependency: parity UNPINNED for its
conventions, restated from the published algorithm (Zhou et al. 2019, "On the Continuity of Rotation Representations in Neural
Networks") as recalled from kaolin/render/camera/extrinsics_backends.py and extrinsics.py:
  params[c] = (a1, a2, t);  b1 = a1/|a1|,  b2 = normalize(a2 - (b1.a2) b1),  b3 = b1 x b2;  R = rows (b1, b2, b3);
  view matrix V = [R | t] (world -> camera);  inv_transform_rays:  o_w = R^T (o_c - t),  d_w = R^T d_c.
Plain differentiable torch, so autograd provides the backward the CUDA kernel is checked against.
"""
import torch


def rot6d_to_matrix(p6):
    """[C,6] -> R [C,3,3] with rows b1, b2, b3 (Gram-Schmidt)."""
    a1, a2 = p6[:, 0:3], p6[:, 3:6]
    b1 = a1 / a1.norm(dim=-1, keepdim=True)
    u2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = u2 / u2.norm(dim=-1, keepdim=True)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3], dim=1)


def params_from_view_matrix(V):
    """[C,4,4] world->camera matrices -> [C,9] parameters (first two rows of R, then t)."""
    return torch.cat([V[:, 0, :3], V[:, 1, :3], V[:, :3, 3]], dim=1)


def transform_rays(params, cam_idx, base_o, base_d):
    """params [n_cam,9], cam_idx int64 [C], base_o / base_d [C*B,3] grouped by camera -> world-space (o, d normalised)."""
    C = cam_idx.shape[0]
    p = params[cam_idx]
    R = rot6d_to_matrix(p[:, :6])                # [C,3,3]
    t = p[:, 6:9]
    o = base_o.reshape(C, -1, 3) - t[:, None, :]
    d = base_d.reshape(C, -1, 3)
    ow = torch.einsum('cij,cbi->cbj', R, o)      # R^T (o - t)
    dw = torch.einsum('cij,cbi->cbj', R, d)
    dw = dw / torch.linalg.norm(dw, dim=-1, keepdim=True)
    return ow.reshape(-1, 3), dw.reshape(-1, 3)
