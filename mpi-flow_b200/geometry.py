"""Camera helpers -- drop-in for MPI-Flow's geometry.py (monodepth2-style API: BackprojectDepth,
Project3D, transformation_from_parameters, get_translation_matrix, rot_from_axisangle;
reference geometry.py:14-153).  HW-sized host-side torch code, written fresh; the arithmetic
order is kept so index maps for the forward splat come out bit-identical.
"""
import torch
import torch.nn as nn

__all__ = ["BackprojectDepth", "Project3D", "transformation_from_parameters"]


class BackprojectDepth(nn.Module):
    """depth Bx1xHxW + inverse intrinsics -> homogeneous camera points Bx4xHW; reference :17-49."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32),
                                torch.arange(width, dtype=torch.float32), indexing="ij")
        self.id_coords = nn.Parameter(torch.stack((xs, ys), 0), requires_grad=False)
        self.ones = nn.Parameter(torch.ones(batch_size, 1, height * width), requires_grad=False)
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1)), 0).unsqueeze(0).repeat(batch_size, 1, 1)
        self.pix_coords = nn.Parameter(torch.cat((pix, self.ones), 1), requires_grad=False)

    def forward(self, depth, inv_K):
        rays = torch.matmul(inv_K[:, :3, :3], self.pix_coords)
        pts = depth.view(self.batch_size, 1, -1) * rays
        return torch.cat((pts, self.ones), 1)


class Project3D(nn.Module):
    """camera points Bx4xHW -> normalised pixel grid BxHxWx2 in [-1,1] and depth Bx1xHW; :52-76."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T, T2=None):
        if T2 is not None:
            T = torch.matmul(T, torch.inverse(T2))
        cam = torch.matmul(torch.matmul(K, T)[:, :3, :], points)
        z = cam[:, 2, :].unsqueeze(1)
        pix = (cam[:, :2, :] / (z + self.eps)).view(self.batch_size, 2, self.height, self.width)
        pix = pix.permute(0, 2, 3, 1)
        pix[..., 0] /= self.width - 1
        pix[..., 1] /= self.height - 1
        return (pix - 0.5) * 2, z


def transformation_from_parameters(axisangle, translation, invert=False):
    """(axis-angle Bx1x3, translation Bx1x3) -> 4x4: T.R, or R^T.T(-t) when invert; :79-95."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def get_translation_matrix(translation_vector):
    """Bx(1x)3 -> Bx4x4 homogeneous translation; reference :98-111."""
    B = translation_vector.shape[0]
    T = torch.eye(4, device=translation_vector.device).repeat(B, 1, 1)
    T[:, :3, 3] = translation_vector.contiguous().view(B, 3)
    return T


def rot_from_axisangle(vec):
    """Rodrigues formula, vec Bx1x3 -> Bx4x4; reference :114-153."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rows = [x * xC + ca, xyC - zs, zxC + ys,
            xyC + zs, y * yC + ca, yzC - xs,
            zxC - ys, yzC + xs, z * zC + ca]
    rot = torch.zeros((vec.shape[0], 4, 4), device=vec.device)
    rot[:, :3, :3] = torch.stack([r.reshape(-1) for r in rows], 1).view(-1, 3, 3)
    rot[:, 3, 3] = 1
    return rot
