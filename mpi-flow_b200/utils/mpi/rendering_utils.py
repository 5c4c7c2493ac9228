"""Drop-in for MPI-Flow's utils/mpi/rendering_utils.py (same five functions, :4,26,46,69,90).

transform_G_xyz is on the renderer's path and runs on the sm_100a library; the four MINE training
helpers are never called by the generator and stay thin torch code, written fresh.
"""
import torch

import _native as nat


def transform_G_xyz(G, xyz, is_return_homo=False):
    """(G . [xyz; 1]) for G Bx4x4 (or 4x4) and xyz Bx3xN (or 3xN); reference :4-23."""
    assert G.dim() == xyz.dim()
    if G.dim() == 2:
        G, xyz = G.unsqueeze(0), xyz.unsqueeze(0)
    B, _, N = xyz.shape
    src = nat.dev(xyz, "xyz").contiguous()
    g = nat.dev(G, "G").contiguous()
    out = torch.empty_like(src)
    nat.check(nat.lib().mpiflow_tgt_xyz_f32(nat.ptr(src), nat.ptr(g), B, 1, 1, N, nat.ptr(out), nat.stream(src)))
    out = out.to(xyz.dtype)
    if is_return_homo:
        # fourth row of a rigid G is (0,0,0,1): the homogeneous coordinate stays 1 for such G; computed
        # generally for API parity
        w = (g[:, 3:4, 0:3] @ src + g[:, 3:4, 3:4]).to(xyz.dtype)
        return torch.cat((out, w), dim=1)
    return out


def gather_pixel_by_pxpy(img, pxpy):
    """Nearest-pixel lookup: img BxCxHxW at pxpy Bx2xN (x, y) -> BxCxN; reference :26-43."""
    B, C, H, W = img.shape
    with torch.no_grad():
        ij = torch.round(pxpy).long() if pxpy.is_floating_point() else pxpy.long()
        flat = ij[:, 0].clamp(0, W - 1) + W * ij[:, 1].clamp(0, H - 1)
    return img.reshape(B, C, H * W).gather(2, flat.unsqueeze(1).expand(B, C, -1))


def uniformly_sample_disparity_from_bins(batch_size, disparity_np, device):
    """One uniform sample inside each disparity bin (edges decreasing); reference :46-66."""
    assert disparity_np[0] > disparity_np[-1]
    edges = torch.as_tensor(disparity_np, dtype=torch.float32, device=device)
    lo, width = edges[:-1], edges[1:] - edges[:-1]
    u = torch.rand((batch_size, width.numel()), dtype=torch.float32, device=device)
    return lo.unsqueeze(0) + width.unsqueeze(0) * u


def uniformly_sample_disparity_from_linspace_bins(batch_size, num_bins, start, end, device):
    """Same with num_bins equal bins between start > end; reference :69-87."""
    assert start > end
    edges = torch.linspace(start, end, num_bins + 1, dtype=torch.float32, device=device)
    u = torch.rand((batch_size, num_bins), dtype=torch.float32, device=device)
    return edges[:-1].unsqueeze(0) + (edges[1] - edges[0]) * u


def sample_pdf(values, weights, N_samples):
    """Inverse-CDF sampling of Bx1xNxS (values, weights) -> Bx1xNxN_samples; reference :90-139."""
    B, _, N, S = weights.shape
    assert values.shape == (B, 1, N, S)
    mids = 0.5 * (values[..., 1:] + values[..., :-1])
    edges = torch.cat((values[..., :1], mids, values[..., -1:]), dim=3)             # S+1 bin edges
    pdf = weights / (weights.sum(dim=3, keepdim=True) + 1e-5)
    cdf = torch.cat((torch.zeros_like(pdf[..., :1]), pdf.cumsum(dim=3)), dim=3)      # S+1
    u = torch.rand((B, 1, N, N_samples), dtype=weights.dtype, device=weights.device)
    hi = torch.searchsorted(cdf, u, right=True)
    lo, hi = (hi - 1).clamp(min=0), hi.clamp(max=S)
    cdf_lo, cdf_hi = cdf.gather(3, lo), cdf.gather(3, hi)
    bin_lo, bin_hi = edges.gather(3, lo), edges.gather(3, hi)
    span = cdf_hi - cdf_lo
    t = torch.where(span <= 1e-4, torch.full_like(u, 0.5), (u - cdf_lo) / span.clamp(min=1e-5))
    return bin_lo + t * (bin_hi - bin_lo)
