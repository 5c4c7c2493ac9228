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
        col = torch.arange(width, dtype=torch.float32).repeat(height)                 # x of pixel k = k % width
        row = torch.arange(height, dtype=torch.float32).repeat_interleave(width)      # y of pixel k = k // width
        frozen = lambda t: nn.Parameter(t, requires_grad=False)
        self.id_coords = frozen(torch.stack((col.view(height, width), row.view(height, width)), 0))
        self.ones = frozen(torch.ones(batch_size, 1, height * width))
        homog = torch.stack((col, row, torch.ones(height * width)), 0)                # 3 x HW: (x, y, 1)
        self.pix_coords = frozen(homog.unsqueeze(0).repeat(batch_size, 1, 1))

    def forward(self, depth, inv_K):
        rays = torch.matmul(inv_K[:, :3, :3], self.pix_coords)
        scaled = depth.view(self.batch_size, 1, -1) * rays
        return torch.cat((scaled, self.ones), 1)


class Project3D(nn.Module):
    """camera points Bx4xHW -> normalised pixel grid BxHxWx2 in [-1,1] and depth Bx1xHW; :52-76."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T, T2=None):
        pose = T if T2 is None else torch.matmul(T, torch.inverse(T2))
        projected = torch.matmul(torch.matmul(K, pose)[:, :3, :], points)             # B x 3 x HW
        z = projected[:, 2:3, :]
        uv = (projected[:, :2, :] / (z + self.eps)).view(self.batch_size, 2, self.height, self.width).permute(0, 2, 3, 1)
        uv[..., 0] /= self.width - 1
        uv[..., 1] /= self.height - 1
        return (uv - 0.5) * 2, z


def transformation_from_parameters(axisangle, translation, invert=False):
    """(axis-angle Bx1x3, translation Bx1x3) -> 4x4: T.R, or R^T.T(-t) when invert; :79-95."""
    rotation = rot_from_axisangle(axisangle)
    if invert:
        return torch.matmul(rotation.transpose(1, 2), get_translation_matrix(translation * -1))
    return torch.matmul(get_translation_matrix(translation), rotation)


def get_translation_matrix(translation_vector):
    """Bx(1x)3 -> Bx4x4 homogeneous translation; reference :98-111."""
    n = translation_vector.shape[0]
    out = torch.eye(4, device=translation_vector.device).repeat(n, 1, 1)
    out[:, :3, 3] = translation_vector.contiguous().view(n, 3)
    return out


def rot_from_axisangle(vec):
    """Rodrigues formula, vec Bx1x3 -> Bx4x4; reference :114-153 (same products in the same order)."""
    theta = torch.norm(vec, 2, 2, True)
    unit = vec / (theta + 1e-7)
    ca, sa = torch.cos(theta), torch.sin(theta)
    C = 1 - ca
    x, y, z = (unit[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    nine = [x * xC + ca, xyC - zs, zxC + ys,
            xyC + zs, y * yC + ca, yzC - xs,
            zxC - ys, yzC + xs, z * zC + ca]
    out = torch.zeros((vec.shape[0], 4, 4), device=vec.device)
    out[:, :3, :3] = torch.stack([e.reshape(-1) for e in nine], 1).view(-1, 3, 3)
    out[:, 3, 3] = 1
    return out
