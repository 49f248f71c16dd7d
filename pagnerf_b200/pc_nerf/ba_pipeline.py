"""BAPipeline: nef + tracer + a trainable camera-pose table (bundle adjustment inside the training step).

Drop-in for the reference class (pc_nerf/ba_pipeline.py:10-92): same constructor (`nef, cameras, tracer, anchor_frame_idxs,
pose_opt_only_frame_idxs`), the `camera_extrinsics` parameter [n_cameras, 9] in kaolin's 'matrix_6dof_rotation' layout
(6-D rotation + translation, :44-51), the anchor-frame gradient mask (:53-62), `forward(..., cam_ids=)` (:64-75) and
`transform_rays(base_rays, cam_ids)` (:85-92).  The transform and its backward are csrc/pose.cu (ops.PoseTransformFn);
kaolin's `Camera` class itself is not needed: `cameras` may be

  * a tensor [n_cameras, 4, 4] of view (world -> camera) matrices, or [n_cameras, 9] of parameters;
  * a dict {cam_id: camera} / list of cameras / one batched camera, where a camera is anything exposing
    `.extrinsics.view_matrix()` (kaolin's API) or `.view_matrix()`; `.near` / `.far` are picked up when present.
"""
import torch
import torch.nn as nn

from .. import ops
from ..wisp_compat import Rays


def _view_matrices(cam):
    ext = getattr(cam, 'extrinsics', cam)
    V = ext.view_matrix() if hasattr(ext, 'view_matrix') else ext
    V = torch.as_tensor(V, dtype=torch.float32)
    return V.reshape(-1, 4, 4)


class _Extrinsics:
    """What the reference's trainer touches on `pipeline.cameras.extrinsics` (pc_nerf/trainer.py:297,308)."""

    def __init__(self, owner):
        self._owner = owner

    def parameters(self):
        return self._owner.camera_extrinsics

    def __len__(self):
        return self._owner.camera_extrinsics.shape[0]


class _Cameras:
    def __init__(self, owner, near, far):
        self.extrinsics = _Extrinsics(owner)
        self.near, self.far = near, far

    def __len__(self):
        return len(self.extrinsics)


class BAPipeline(nn.Module):
    def __init__(self, nef, cameras, tracer=None, anchor_frame_idxs=(), pose_opt_only_frame_idxs=(), near=0.0, far=6.0):
        super().__init__()
        self.nef, self.tracer = nef, tracer
        self.cam_id_to_idx = None
        if isinstance(cameras, dict):
            self.cam_id_to_idx = {cam_id: idx for idx, cam_id in enumerate(cameras.keys())}
            cams = list(cameras.values())
        elif isinstance(cameras, (tuple, list)):
            cams = list(cameras)
        else:
            cams = [cameras]
        if torch.is_tensor(cams[0]) and cams[0].shape[-1] == 9 and cams[0].dim() == 2:
            params = cams[0].to(torch.float32).clone()
        else:
            V = torch.cat([_view_matrices(c) for c in cams], dim=0)
            params = torch.cat([V[:, 0, :3], V[:, 1, :3], V[:, :3, 3]], dim=1)      # first two rows of R, then t
        if params.shape[0] <= 1 and not isinstance(cameras, (dict, tuple, list)):
            raise AssertionError('Tried to create a camera database module with a single camera extrinsics, but needs more than one')
        c0 = cams[0]
        near = float(getattr(c0, 'near', near)) if not torch.is_tensor(c0) else near
        far = float(getattr(c0, 'far', far)) if not torch.is_tensor(c0) else far
        self.anchor_frame_idxs = list(anchor_frame_idxs)
        self.pose_opt_only_frame_idxs = list(pose_opt_only_frame_idxs)
        self.camera_extrinsics = nn.Parameter(params)
        self.cameras = _Cameras(self, near, far)
        self._mask_hook = None

    def to(self, *args, **kwargs):
        self = super().to(*args, **kwargs)
        if len(self.anchor_frame_idxs) > 0:      # anchor frames keep their pose: zero their gradient rows (:53-62)
            if self._mask_hook is not None:
                self._mask_hook.remove()
            grad_mask = torch.ones_like(self.camera_extrinsics)
            grad_mask[self.anchor_frame_idxs] = 0.0
            self._mask_hook = self.camera_extrinsics.register_hook(lambda grad: grad * grad_mask)
        return self

    def forward(self, *args, cam_ids=None, **kwargs):
        """Transform the base rays with the requested camera poses, then trace (or evaluate the field)."""
        if isinstance(cam_ids, (tuple, list, torch.Tensor)):
            kwargs['rays'] = self.transform_rays(kwargs['rays'], cam_ids)
        if self.tracer is not None:
            return self.tracer(self.nef, *args, **kwargs)
        return self.nef(*args, **kwargs)

    def get_camera_indices(self, cam_ids):
        assert isinstance(cam_ids, (tuple, list, torch.Tensor))
        if isinstance(cam_ids, (tuple, list)):
            if self.cam_id_to_idx is not None:
                cam_ids = torch.tensor([self.cam_id_to_idx[i] for i in cam_ids], dtype=torch.long)
            else:
                cam_ids = torch.tensor([int(i) for i in cam_ids], dtype=torch.long)
        assert cam_ids.nelement() > 0
        return cam_ids.to(self.camera_extrinsics.device)

    def transform_rays(self, base_rays, cam_ids):
        idx = self.get_camera_indices(cam_ids)
        o, d = ops.pose_transform(self.camera_extrinsics, idx, base_rays.origins.reshape(-1, 3), base_rays.dirs.reshape(-1, 3))
        return Rays(origins=o, dirs=d, dist_min=self.cameras.near, dist_max=self.cameras.far)
