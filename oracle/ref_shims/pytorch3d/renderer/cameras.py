"""Restatement of the pytorch3d PerspectiveCameras arithmetic the reference touches (pytorch3d is un-vendored).

Row-vector convention: X_view = X_world @ R + T.  NDC: x = fx X/Z + px, y = fy Y/Z + py, z = 1/Z.
"""
import torch


class CamerasBase:
    pass


class _W2V:
    def __init__(self, R, T):
        self.R, self.T = R, T

    def get_matrix(self):
        n = self.R.shape[0]
        M = torch.zeros(n, 4, 4, dtype=self.R.dtype, device=self.R.device)
        M[:, :3, :3] = self.R
        M[:, 3, :3] = self.T
        M[:, 3, 3] = 1
        return M

    def inverse(self):
        Rt = self.R.transpose(1, 2)
        return _W2V(Rt, -torch.einsum("bi,bij->bj", self.T, Rt))

    def compose(self, other):
        # apply self first, then other:  M = M_self @ M_other
        R = self.R @ other.R
        T = torch.einsum("bi,bij->bj", self.T.expand(R.shape[0], -1), other.R) + other.T
        return _W2V(R, T)


class PerspectiveCameras(CamerasBase):
    def __init__(self, R=None, T=None, focal_length=1.0, principal_point=None, image_size=None, device="cpu", **kw):
        if R is None:
            R = torch.eye(3)[None]
        if T is None:
            T = torch.zeros(1, 3)
        n = max(R.shape[0], T.shape[0])
        self.R = R.float().expand(n, 3, 3).to(device)
        self.T = T.float().expand(n, 3).to(device)
        if not torch.is_tensor(focal_length):
            focal_length = torch.full((n, 2), float(focal_length))
        if principal_point is None:
            principal_point = torch.zeros(n, 2)
        self.focal_length = focal_length.float().to(device)
        self.principal_point = principal_point.float().to(device)
        self.image_size = image_size
        self.device = torch.device(device) if not isinstance(device, torch.device) else device

    def __len__(self):
        return self.R.shape[0]

    def to(self, device):
        return PerspectiveCameras(R=self.R, T=self.T, focal_length=self.focal_length,
                                  principal_point=self.principal_point, image_size=self.image_size, device=device)

    def get_camera_center(self):
        return -torch.einsum("bj,bij->bi", self.T, self.R)

    def get_world_to_view_transform(self):
        return _W2V(self.R, self.T)

    def transform_points_ndc(self, points):
        # points (1 or B, P, 3) -> (B, P, 3)
        v = points @ self.R + self.T[:, None, :]
        x = self.focal_length[:, None, 0] * v[..., 0] / v[..., 2] + self.principal_point[:, None, 0]
        y = self.focal_length[:, None, 1] * v[..., 1] / v[..., 2] + self.principal_point[:, None, 1]
        return torch.stack([x, y, 1.0 / v[..., 2]], dim=-1)

    def unproject_points(self, xy_depth, world_coordinates=True, from_ndc=False, **kw):
        d = xy_depth[..., 2]
        X = (xy_depth[..., 0] - self.principal_point[:, None, 0]) * d / self.focal_length[:, None, 0]
        Y = (xy_depth[..., 1] - self.principal_point[:, None, 1]) * d / self.focal_length[:, None, 1]
        v = torch.stack([X, Y, d], dim=-1)
        if not world_coordinates:
            return v
        return (v - self.T[:, None, :]) @ self.R.transpose(1, 2)


def look_at_view_transform(*a, **k):  # import-only stub
    raise NotImplementedError
