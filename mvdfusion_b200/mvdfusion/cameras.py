"""Minimal camera container with the four tensor fields the hot path reads from pytorch3d's
PerspectiveCameras (`.R .T .focal_length .principal_point`); any object exposing those works."""
import torch


class PerspectiveCameras:
    def __init__(self, R, T, focal_length, principal_point=None, device=None, image_size=None):
        n = R.shape[0]
        dev = device if device is not None else R.device
        self.R = R.float().to(dev)
        self.T = T.float().expand(n, 3).to(dev)
        f = focal_length if torch.is_tensor(focal_length) else torch.full((n, 2), float(focal_length))
        self.focal_length = f.float().reshape(-1, f.shape[-1]).expand(n, 2).to(dev)
        pp = principal_point if principal_point is not None else torch.zeros(n, 2)
        self.principal_point = pp.float().expand(n, 2).to(dev)
        self.image_size = image_size
        self.device = torch.device(dev)

    def __len__(self):
        return self.R.shape[0]

    def to(self, device):
        return PerspectiveCameras(self.R, self.T, self.focal_length, self.principal_point, device=device)

    def __getitem__(self, idx):
        if isinstance(idx, int):
            idx = [idx]
        return PerspectiveCameras(self.R[idx], self.T[idx], self.focal_length[idx], self.principal_point[idx])

    def get_camera_center(self):
        return -torch.einsum("bj,bij->bi", self.T, self.R)


def relative_cameras(cams, query_idx):
    """utils/camera_utils.py:58-115 (center_at_origin=False): world frame rotated so that the query camera has R = I;
    R' = R_q^T R, T' = T."""
    Rq = cams.R[query_idx][0] if torch.is_tensor(query_idx) or isinstance(query_idx, (list, tuple)) else cams.R[query_idx]
    R = torch.einsum("ji,bjk->bik", Rq, cams.R)
    return PerspectiveCameras(R, cams.T, cams.focal_length, cams.principal_point)
