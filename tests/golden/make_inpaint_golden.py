"""Fixtures of cv2.inpaint(img, mask, radius, cv2.INPAINT_NS) -- the hole filling of the reference's sample assembly
(utils/utils.py:284-286; OpenCV is an un-vendored dependency of the reference, cv2 4.13 in this image).

    python tests/golden/make_inpaint_golden.py        -> tests/golden/inpaint.npz

Seeded images (white noise, two smoothness levels) with line / disc / noise holes, several of them touching the image
border; radius 3 (the reference's) and a few others.  The outputs are cv2's, stored next to the inputs, so the tests
do not depend on the cv2 build of the machine they run on.
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(rng, H, W, smooth, kind, border):
    img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
    if smooth > 0:
        img = cv2.GaussianBlur(img, (0, 0), smooth)
    mask = np.zeros((H, W), np.uint8)
    if kind == "noise":
        mask = (rng.random((H, W)) < 0.25).astype(np.uint8)
    else:
        for _ in range(int(rng.integers(4, 14))):
            x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
            if rng.random() < 0.6:
                x2 = int(np.clip(x + rng.integers(-30, 30), 0, W - 1))
                y2 = int(np.clip(y + rng.integers(-30, 30), 0, H - 1))
                cv2.line(mask, (x, y), (x2, y2), 1, int(rng.integers(1, 7)))
            else:
                cv2.circle(mask, (x, y), int(rng.integers(1, 9)), 1, -1)
    if border:
        mask[0, : W // 3] = 1
        mask[H // 3: H // 2, 0] = 1
        mask[-1, -W // 4:] = 1
        mask[-H // 4: -2, -1] = 1
        mask[:3, -3:] = 1
    else:
        mask[:5] = 0; mask[-5:] = 0; mask[:, :5] = 0; mask[:, -5:] = 0
    return img, mask


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    n = 0
    for H, W in ((48, 64), (37, 91), (96, 160)):
        for smooth in (0, 1.5, 4):
            for kind in ("shapes", "noise"):
                for border in (False, True):
                    radius = 3 if n % 4 else int(rng.integers(1, 6))
                    img, mask = make_case(rng, H, W, smooth, kind, border)
                    res = cv2.inpaint(img, mask, radius, cv2.INPAINT_NS)
                    out[f"img{n}"], out[f"mask{n}"], out[f"out{n}"] = img, mask, res
                    out[f"radius{n}"] = np.int32(radius)
                    n += 1
    # a few small cases on which the double-precision reading of the direction term (oracle variant 2) differs from
    # cv2: they are rare (about 2 per 10 000 filled values), so they are searched for rather than hoped for
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import inpaint_oracle as io
    found = 0
    while found < 4:
        img, mask = make_case(rng, 40, 56, 0, "noise" if found % 2 else "shapes", False)
        radius = int(rng.integers(2, 6))
        res = cv2.inpaint(img, mask, radius, cv2.INPAINT_NS)
        if not np.array_equal(io.inpaint_ns(img, mask, radius, 2), res):
            out[f"img{n}"], out[f"mask{n}"], out[f"out{n}"], out[f"radius{n}"] = img, mask, res, np.int32(radius)
            n += 1
            found += 1
    out["count"] = np.int32(n)
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(HERE, "inpaint.npz")
    np.savez_compressed(path, **out)
    print(path, n, "cases", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
