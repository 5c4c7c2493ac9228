"""Synthetic inputs for the MPI renderer hot path.

numpy only (no CUDA dependency): used by tests/, bench.py and tests/golden/make_golden.py so that
fixtures can be regenerated from a seed instead of being stored.

Conventions follow the reference driver:
  K            gen_3dphoto_dynamic_v2.py:42-49   [[0.58W,0,0.5W],[0,0.58H,0.5H],[0,0,1]]
  disparity    model/AdaMPI.py:64-68             linspace(1, 0.001, S+2)[1:-1]
  MPI          model/CPN/decoder.py:168-172      rgb in (0,1), sigma = relu(.) + 1e-4
  pose         utils/utils.py:121-156            generate_random_pose(ext_cz=0.15)
  seed         gen_3dphoto_dynamic_v2.py:24      114514
"""
import math

import numpy as np

SEED = 114514


def intrinsics(H, W, B=1):
    K = np.array([[0.58 * W, 0, 0.5 * W], [0, 0.58 * H, 0.5 * H], [0, 0, 1]], np.float32)
    return np.repeat(K[None], B, 0)


def intrinsics_inv(K):
    """fp64 inverse rounded to fp32, as utils/utils.py:186-187."""
    return np.linalg.inv(K.astype(np.float64)).astype(np.float32)


def plane_disparities(S, B=1):
    d = np.linspace(1.0, 0.001, S + 2, dtype=np.float64)[1:-1].astype(np.float32)
    return np.repeat(d[None], B, 0)


def _box_blur(a, k):
    """separable k x k box filter with edge replication over the last two axes"""
    r = k // 2
    for ax in (-1, -2):
        pad = [(0, 0)] * a.ndim
        pad[ax] = (r, r)
        p = np.pad(a, pad, mode="edge")
        c = np.cumsum(p, axis=ax, dtype=np.float64)
        c = np.concatenate([np.zeros_like(np.take(c, [0], axis=ax)), c], axis=ax)
        n = a.shape[ax]
        hi = np.take(c, np.arange(k, k + n), axis=ax)
        lo = np.take(c, np.arange(0, n), axis=ax)
        a = ((hi - lo) / k).astype(np.float32)
    return a


def mpi(B, S, H, W, kind="white", seed=SEED):
    """[B,S,4,H,W] fp32: rgb uniform(0,1), sigma = relu(N(0,1)) + 1e-4; 'smooth' = 9x9 box filtered."""
    rng = np.random.default_rng(seed)
    rgb = rng.random((B, S, 3, H, W), dtype=np.float32)
    sig = rng.standard_normal((B, S, 1, H, W), dtype=np.float32)
    if kind == "smooth":
        rgb = _box_blur(rgb, 9)
        sig = _box_blur(sig, 9) * 4.0
    elif kind != "white":
        raise ValueError(kind)
    sig = np.maximum(sig, 0) + np.float32(1e-4)
    return np.ascontiguousarray(np.concatenate([rgb, sig], 2).astype(np.float32))


def obj_mask(B, H, W, seed=SEED):
    """[B,1,H,W] soft-edged blobs in [0,1] (a coarse 0/1 grid, bilinearly upsampled)."""
    rng = np.random.default_rng(seed + 1)
    gh, gw = max(2, H // 16), max(2, W // 16)
    coarse = (rng.random((B, gh, gw)) > 0.6).astype(np.float32)
    ys = (np.arange(H) + 0.5) * gh / H - 0.5
    xs = (np.arange(W) + 0.5) * gw / W - 0.5
    y0 = np.clip(np.floor(ys), 0, gh - 1).astype(int); y1 = np.clip(y0 + 1, 0, gh - 1)
    x0 = np.clip(np.floor(xs), 0, gw - 1).astype(int); x1 = np.clip(x0 + 1, 0, gw - 1)
    fy = np.clip(ys - y0, 0, 1)[None, :, None].astype(np.float32)
    fx = np.clip(xs - x0, 0, 1)[None, None, :].astype(np.float32)
    top = coarse[:, y0][:, :, x0] * (1 - fx) + coarse[:, y0][:, :, x1] * fx
    bot = coarse[:, y1][:, :, x0] * (1 - fx) + coarse[:, y1][:, :, x1] * fx
    return np.ascontiguousarray((top * (1 - fy) + bot * fy)[:, None].astype(np.float32))


def pose_from_parameters(axisangle, translation):
    """4x4 = T(translation) @ R(axisangle), fp32 (same maths as geometry.py:79-153, non-inverted)."""
    v = np.asarray(axisangle, np.float32)
    ang = np.float32(np.sqrt(np.sum(v * v, dtype=np.float32)))
    ax = v / (ang + np.float32(1e-7))
    ca, sa = np.float32(math.cos(ang)), np.float32(math.sin(ang))
    C = np.float32(1) - ca
    x, y, z = ax
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = np.array([[x * x * C + ca, x * y * C - z * sa, z * x * C + y * sa],
                          [x * y * C + z * sa, y * y * C + ca, y * z * C - x * sa],
                          [z * x * C - y * sa, y * z * C + x * sa, z * z * C + ca]], np.float32)
    M[:3, 3] = np.asarray(translation, np.float32)
    return M


REGRESSION_POSE = ((0.01, -0.02, 0.005), (0.1, -0.05, -0.15))


def random_pose(rng, ext_cz=0.15, object_motion=False):
    """Distribution of utils/utils.py:121-156 (camera pose, or the 'object' pose with base 0)."""
    base = 0.0 if object_motion else 0.1
    sgn = lambda: (-1.0) ** int(rng.integers(2))
    scx, scy, scz = sgn(), sgn(), sgn()
    if object_motion:
        scx, scy, scz = scx * 0.5, scy * 0.5, scz * 0.5
    else:
        scz = -1.0
    t = ((rng.random() * 0.1 + base) * scx, (rng.random() * 0.1 + base) * scy,
         (rng.random() * ext_cz + base) * scz)
    a = tuple(rng.random() * math.pi / 36.0 * sgn() * 0.4 for _ in range(3))
    return pose_from_parameters(a, t)


def poses(B, kind="random", seed=SEED):
    rng = np.random.default_rng(seed + 2)
    if kind == "identity":
        return np.repeat(np.eye(4, dtype=np.float32)[None], B, 0)
    if kind == "regression":
        return np.repeat(pose_from_parameters(*REGRESSION_POSE)[None], B, 0)
    if kind == "object":
        return np.stack([random_pose(rng, object_motion=True) for _ in range(B)])
    return np.stack([random_pose(rng) for _ in range(B)])


def frame_inputs(B, S, H, W, kind="white", pose="random", seed=SEED):
    """Everything one render_tgt_rgb_depth-equivalent needs, as a dict of fp32 numpy arrays."""
    K = intrinsics(H, W, B)
    return dict(mpi=mpi(B, S, H, W, kind, seed), disparity=plane_disparities(S, B), G=poses(B, pose, seed),
                K=K, Kinv=intrinsics_inv(K), obj_mask=obj_mask(B, H, W, seed))


def splat_inputs(h, w, seed=SEED, collisions="dense"):
    """Random forward-splat problem: u8 image, clamped int64 index maps with many collisions,
    fp32 z with ties and exact 1000.0 values (warping.c's sentinel)."""
    rng = np.random.default_rng(seed + 3)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    if collisions == "dense":
        idx = np.clip(xx + rng.integers(-3, 4, (h, w)), 0, w - 1)
        idy = np.clip(yy + rng.integers(-3, 4, (h, w)), 0, h - 1)
    elif collisions == "pile":      # every source pixel lands on a handful of targets
        idx = rng.integers(0, max(1, w // 8), (h, w))
        idy = rng.integers(0, max(1, h // 8), (h, w))
    else:                           # identity map: no collisions at all
        idx, idy = xx, yy
    z = rng.choice(np.array([0.5, 1.0, 2.0, 2.0, 7.25, 1000.0, 1000.0, 1500.0], np.float32), (h, w))
    z = np.where(rng.random((h, w)) < 0.5, z, rng.random((h, w), dtype=np.float32) * 50).astype(np.float32)
    return src, idx.astype(np.int64), idy.astype(np.int64), z
