"""CPU restatement of pytorch3d 0.6.2's naive mesh rasteriser + attribute interpolation -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference (warpback/utils.py:56-63) calls pytorch3d.renderer.mesh.rasterize_meshes(mesh,
(h, w), faces_per_pixel=1, blur_radius=1e-6) and pytorch3d.ops.interpolate_face_attributes; pytorch3d 0.6.2
(README.md:30) is an un-vendored dependency that is not installed in this image, and the reference holds no test or
fixture for this path.  This file restates the published algorithm
(pytorch3d/csrc/rasterize_meshes/rasterize_meshes.cu RasterizeMeshesNaiveCudaKernel, csrc/utils/geometry_utils.cuh,
renderer/mesh/rasterize_meshes.py): per pixel, over all faces in index order,
  * pixel centre in NDC, both image axes reversed (+X left, +Y up); the shorter side spans [-1, 1];
  * skip faces with min z < 1e-8, |area| <= 1e-8, or whose box (grown by sqrt(blur_radius)) misses the pixel;
  * barycentrics from edge functions with area + 1e-8; clipped to [0,1] and renormalised when blur_radius > 0
    (clip_barycentric_coords defaults to blur_radius > 0);
  * pz = bary . z, skipped when < 0; outside pixels count only if their squared distance to the triangle is < blur;
  * the face with the smallest pz wins, the first one on ties.
numpy float32 arithmetic in the kernel's order; loops over pixels: small cases only."""
import numpy as np

F32 = np.float32
EPS = F32(1e-8)


def _pix_to_ndc(i, s1, s2):
    rng = F32(2.0) * F32(s1) / F32(s2) if s1 > s2 else F32(2.0)
    off = rng * F32(0.5)
    return -off + (rng * F32(i) + off) / F32(s1)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist2(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.clip((bax * (px - ax) + bay * (py - ay)) / l2, F32(0), F32(1))
    qx, qy = ax + t * bax - px, ay + t * bay - py
    d = qx * qx + qy * qy
    return np.where(l2 <= EPS, (px - bx) ** 2 + (py - by) ** 2, d).astype(F32)


def rasterize(verts, faces, attrs, H, W, blur_radius=1e-6):
    """verts [V,3] f32 (NDC x, y, depth z), faces [F,3] int, attrs [V,C] -> dict(pix_to_face [H,W] (-1 = none),
    zbuf [H,W], bary [H,W,3], out [C,H,W])"""
    verts = np.asarray(verts, F32)
    attrs = np.asarray(attrs, F32)
    faces = np.asarray(faces)
    blur = F32(blur_radius)
    br = np.sqrt(blur, dtype=F32)
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    x0, y0, z0 = v0[:, 0], v0[:, 1], v0[:, 2]
    x1, y1, z1 = v1[:, 0], v1[:, 1], v1[:, 2]
    x2, y2, z2 = v2[:, 0], v2[:, 1], v2[:, 2]
    zmin = np.minimum(z0, np.minimum(z1, z2))
    area_face = _edge(x0, y0, x1, y1, x2, y2)
    xmin, xmax = np.minimum(x0, np.minimum(x1, x2)) - br, np.maximum(x0, np.maximum(x1, x2)) + br
    ymin, ymax = np.minimum(y0, np.minimum(y1, y2)) - br, np.maximum(y0, np.maximum(y1, y2)) + br
    usable = (zmin >= EPS) & ~((area_face <= EPS) & (area_face >= -EPS))
    area = _edge(x2, y2, x0, y0, x1, y1) + EPS
    C = attrs.shape[1]
    p2f = -np.ones((H, W), np.int64)
    zbuf = -np.ones((H, W), F32)
    bary = -np.ones((H, W, 3), F32)
    out = np.zeros((C, H, W), F32)
    for row in range(H):
        py = _pix_to_ndc(H - 1 - row, H, W)
        for col in range(W):
            px = _pix_to_ndc(W - 1 - col, W, H)
            cand = usable & ~((px > xmax) | (px < xmin) | (py > ymax) | (py < ymin))
            idx = np.nonzero(cand)[0]
            if idx.size == 0:
                continue
            b0 = _edge(px, py, x1[idx], y1[idx], x2[idx], y2[idx]) / area[idx]
            b1 = _edge(px, py, x2[idx], y2[idx], x0[idx], y0[idx]) / area[idx]
            b2 = _edge(px, py, x0[idx], y0[idx], x1[idx], y1[idx]) / area[idx]
            w0, w1, w2 = b0, b1, b2
            if blur > 0:
                w0, w1, w2 = (np.clip(b, F32(0), F32(1)) for b in (b0, b1, b2))
                s = np.maximum(w0 + w1 + w2, F32(1e-5))
                w0, w1, w2 = w0 / s, w1 / s, w2 / s
            pz = (w0 * z0[idx] + w1 * z1[idx] + w2 * z2[idx]).astype(F32)
            inside = (b0 > 0) & (b1 > 0) & (b2 > 0)
            d = np.minimum(_seg_dist2(px, py, x0[idx], y0[idx], x1[idx], y1[idx]),
                           np.minimum(_seg_dist2(px, py, x0[idx], y0[idx], x2[idx], y2[idx]),
                                      _seg_dist2(px, py, x1[idx], y1[idx], x2[idx], y2[idx])))
            ok = (pz >= 0) & (inside | (d < blur))
            if not ok.any():
                continue
            pzk = np.where(ok, pz, np.inf)
            k = int(np.argmin(pzk))            # first minimum = lowest face index on ties
            f = int(idx[k])
            p2f[row, col] = f
            zbuf[row, col] = pz[k]
            bary[row, col] = (w0[k], w1[k], w2[k])
            fv = faces[f]
            out[:, row, col] = w0[k] * attrs[fv[0]] + w1[k] * attrs[fv[1]] + w2[k] * attrs[fv[2]]
    return dict(pix_to_face=p2f, zbuf=zbuf, bary=bary, out=out)
