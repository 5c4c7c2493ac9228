"""Plane-homography sampler -- drop-in for MPI-Flow's utils/mpi/homography_sampler.py.

Same class, method names, argument order and return tuples as the reference
(utils/mpi/homography_sampler.py:30-220); the work is done by libmpiflow_sm100.so:
the plane homographies H = K_t (R - t n^T / -d) K_s^-1 and their fp64 inverses are computed by a
device kernel (no GPU->CPU->GPU round trip as in :15-17), the bilinear/border warp and the
per-plane flows by hand-written sm_100a kernels instead of bmm + ~10 elementwise ops + grid_sample.
"""
import numpy as np
import torch

import _native as nat


def inverse(matrices):
    """Batched inverse of Bx3x3 that raises if the result has NaNs (reference :6-27).

    Stays on the input's device; the reference's retry loop for an old torch bug is not needed."""
    inv = torch.linalg.inv(matrices)
    if torch.isnan(inv).any():
        raise Exception("Matrix inverse contains nan!")
    return inv


class HomographySample:
    def __init__(self, H_tgt, W_tgt, device=None):
        self.device = torch.device("cpu") if device is None else device
        self.Height_tgt = H_tgt
        self.Width_tgt = W_tgt
        self.meshgrid = self.grid_generation(H_tgt, W_tgt, self.device).permute(2, 0, 1).contiguous()
        self.n = self.plane_normal_generation(self.device)

    @staticmethod
    def grid_generation(H, W, device):
        """HxWx3 grid of homogeneous integer pixel coordinates (x, y, 1), no half-pixel offset."""
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=device),
                                torch.arange(W, dtype=torch.float32, device=device), indexing="ij")
        return torch.stack((xs, ys, torch.ones_like(xs)), dim=2)

    @staticmethod
    def plane_normal_generation(device):
        return torch.tensor([0, 0, 1], dtype=torch.float32, device=device)

    @staticmethod
    def euler_to_rotation_matrix(x_angle, y_angle, z_angle, seq='xyz', degrees=False):
        """Rotation taking target-frame points into the source frame (angles negated, :64-78)."""
        from scipy.spatial.transform import Rotation
        rot = Rotation.from_euler(seq, [-x_angle, -y_angle, -z_angle], degrees=degrees)
        return rot.as_matrix().astype(np.float32)

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _homographies(d_src_B, G_tgt_src, K_src_inv, K_tgt, device):
        N = d_src_B.numel()
        G = nat.dev(G_tgt_src.to(device), "G_tgt_src").contiguous()
        Ki = nat.dev(K_src_inv.to(device), "K_src_inv").contiguous()
        Kt = nat.dev(K_tgt.to(device), "K_tgt").contiguous()
        d = nat.dev(d_src_B.to(device), "d_src_B").contiguous()
        assert G.shape == (N, 4, 4) and Ki.shape == (N, 3, 3) and Kt.shape == (N, 3, 3)
        Hts = torch.empty((N, 3, 3), dtype=torch.float32, device=device)
        Hst = torch.empty((N, 3, 3), dtype=torch.float32, device=device)
        status = torch.zeros(1, dtype=torch.int32, device=device)
        nat.check(nat.lib().mpiflow_plane_homographies_f32(nat.ptr(G), nat.ptr(Ki), nat.ptr(Kt), nat.ptr(d), N,
                                                           nat.ptr(Hts), nat.ptr(Hst), nat.ptr(status),
                                                           nat.stream(G)))
        return Hts, Hst, status

    def sample(self, src_BCHW, d_src_B, G_tgt_src, K_src_inv, K_tgt):
        """Warp src (at pose [I|0], plane depth d) into the target view.

        -> (tgt_BCHW, valid_mask_BHW (bool), flowB2A_BHW2); reference :80-158."""
        B, C, H, W = src_BCHW.shape
        assert (H, W) == (self.Height_tgt, self.Width_tgt), "source and target sizes must agree"
        out_dtype = K_tgt.dtype
        src = nat.dev(src_BCHW, "src_BCHW").contiguous()
        _, Hst, status = self._homographies(d_src_B, G_tgt_src, K_src_inv, K_tgt, src.device)
        tgt = torch.empty_like(src)
        valid = torch.empty((B, H, W), dtype=torch.uint8, device=src.device)
        flow = torch.empty((B, H, W, 2), dtype=torch.float32, device=src.device)
        nat.check(nat.lib().mpiflow_homography_sample_f32(nat.ptr(src), nat.ptr(Hst), B, C, H, W, nat.ptr(tgt),
                                                          nat.ptr(valid), nat.ptr(flow), nat.stream(src)))
        if int(status.item()):  # the reference synchronises here as well (CPU inverse)
            raise Exception("Matrix inverse contains nan!")
        return tgt.to(src_BCHW.dtype), valid.bool(), flow.to(out_dtype)

    def sample_inverse(self, src_BCHW, d_src_B, G_tgt_src, K_src_inv, K_tgt):
        """Per-plane forward flow of every source pixel -> flowA2B_BHW2; reference :160-220.
        src_BCHW only supplies the batch size and device."""
        B = src_BCHW.shape[0]
        H, W = self.Height_tgt, self.Width_tgt
        device = src_BCHW.device
        Hts, _, _ = self._homographies(d_src_B, G_tgt_src, K_src_inv, K_tgt, device)
        flow = torch.empty((B, H, W, 2), dtype=torch.float32, device=device)
        nat.check(nat.lib().mpiflow_homography_flow_f32(nat.ptr(Hts), B, H, W, nat.ptr(flow), nat.stream(flow)))
        return flow.to(K_tgt.dtype)
