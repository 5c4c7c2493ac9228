"""Hole filling on the device (csrc/inpaint.cu) against cv2.inpaint(INPAINT_NS) -- reference call
site utils/utils.py:284-286.  Bit-exact: the committed cv2 fixtures, the reference's own `inpainted` image of the sample
fixture, the C oracle at the generator's frame size (1280x384, sliver holes incl. the image border), batches."""
import os

import numpy as np
import pytest
import torch

from oracle import inpaint_oracle as io

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "inpaint.npz"))


def test_device_inpaint_equals_cv2_fixtures(gold):
    from utils import utils as U
    for i in range(int(gold["count"])):
        got = U.inpaint_ns(_dev(gold[f"img{i}"]), _dev(gold[f"mask{i}"]), int(gold[f"radius{i}"]))
        assert np.array_equal(got.cpu().numpy(), gold[f"out{i}"]), f"case {i}"


def test_device_inpaint_reproduces_the_reference_sample(golden_dir):
    from utils import utils as U
    g = np.load(os.path.join(golden_dir, "sample_pair.npz"))
    got = U.inpaint_ns(_dev(g["frame_mix"]), _dev(g["fill_mask"]), 3)
    assert np.array_equal(got.cpu().numpy(), g["inpainted"])


def _sliver_case(rng, H, W, n_lines):
    """a frame-sized image with disocclusion-like slivers, some of them on the image border"""
    import cv2
    img = cv2.GaussianBlur((rng.random((H, W, 3)) * 255).astype(np.uint8), (0, 0), 2.0)
    mask = np.zeros((H, W), np.uint8)
    for _ in range(n_lines):
        x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
        cv2.line(mask, (x, y), (int(np.clip(x + rng.integers(-40, 40), 0, W - 1)),
                                int(np.clip(y + rng.integers(-120, 120), 0, H - 1))), 1, int(rng.integers(1, 6)))
    mask[:, :3] = 1          # the strip a sideways camera move uncovers
    mask[-2:, W // 2:] = 1
    return img, mask


def test_frame_size_batch_equals_oracle_and_cv2():
    """the generator's frame size, a batch of 3 in one launch; the third image has a large blob (long buckets: the
    shared-memory and the global-memory sort) and is also filled in place"""
    cv2 = pytest.importorskip("cv2")
    from utils import utils as U
    rng = np.random.default_rng(2026)
    H, W = 384, 1280
    cases = [_sliver_case(rng, H, W, 60), _sliver_case(rng, H, W, 25), _sliver_case(rng, H, W, 10)]
    cv2.circle(cases[2][1], (600, 190), 150, 1, -1)
    cv2.rectangle(cases[2][1], (900, 0), (1279, 383), 1, -1)      # a front several thousand pixels long
    imgs = np.stack([c[0] for c in cases])
    masks = np.stack([c[1] for c in cases])
    frac = masks.astype(bool).mean(axis=(1, 2))
    assert 0.01 < frac[0] < 0.2 and frac[2] > 0.3
    got = U.inpaint_ns(_dev(imgs), _dev(masks), 3).cpu().numpy()
    for n in range(3):
        want = io.inpaint_ns(imgs[n], masks[n], 3)
        assert np.array_equal(want, cv2.inpaint(imgs[n], masks[n], 3, cv2.INPAINT_NS))
        assert np.array_equal(got[n], want), f"image {n}"
    # one image on its own == its slice of the batch; in place == out of place
    buf = _dev(imgs[2])
    res = U.inpaint_ns(buf, _dev(masks[2]), 3, out=buf)
    assert res.data_ptr() == buf.data_ptr() and np.array_equal(buf.cpu().numpy(), got[2])


def test_empty_and_full_masks():
    from utils import utils as U
    rng = np.random.default_rng(4)
    img = (rng.random((33, 47, 3)) * 255).astype(np.uint8)
    for mask in (np.zeros((33, 47), np.uint8), np.ones((33, 47), np.uint8)):
        assert np.array_equal(U.inpaint_ns(_dev(img), _dev(mask), 3).cpu().numpy(), img)


def test_argument_errors():
    from utils import utils as U
    import _native as nat
    img = torch.zeros((8, 8, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError):
        U.inpaint_ns(img.cpu(), torch.zeros((8, 8), dtype=torch.uint8))
    with pytest.raises(TypeError):
        U.inpaint_ns(img.float(), torch.zeros((8, 8), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        U.inpaint_ns(img, torch.zeros((8, 8), dtype=torch.uint8, device="cuda"), out=torch.empty_like(img))     # in place or nothing
    with pytest.raises(nat.MpiflowError):
        U.inpaint_ns(img, torch.zeros((8, 8), dtype=torch.uint8, device="cuda"), radius=9)   # window larger than the kernel's


def test_holes_at_the_independence_boundary():
    """the device path marches holes that are >= radius + 2 columns (or rows) apart separately: holes placed exactly
    radius + 1, + 2 and + 3 apart, in columns and in rows, on white noise (any leak between two marches shows), for
    radii 1..4; plus 25 random layouts -- all against the pinned C restatement"""
    from utils import utils as U
    rng = np.random.default_rng(99)
    H, W = 72, 120
    cases = []
    for radius in (1, 2, 3, 4):
        for gap in (radius + 1, radius + 2, radius + 3):
            img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
            mask = np.zeros((H, W), np.uint8)
            x = 10
            for _ in range(4):                      # vertical bars `gap` columns apart (first hole column to next)
                mask[8:40, x:x + 3] = 1
                x += 2 + gap                        # last column of a bar = x + 2, next bar starts `gap` columns later
            y = 44
            for _ in range(3):                      # horizontal bars `gap` rows apart, in their own column range
                mask[y:y + 2, 60:110] = 1
                y += 1 + gap
            cases.append((img, mask, radius))
    for _ in range(25):
        img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
        mask = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(3, 12))):
            x0, y0 = int(rng.integers(0, W - 8)), int(rng.integers(0, H - 8))
            mask[y0:y0 + int(rng.integers(1, 8)), x0:x0 + int(rng.integers(1, 8))] = 1
        cases.append((img, mask, int(rng.integers(1, 5))))
    for n, (img, mask, radius) in enumerate(cases):
        got = U.inpaint_ns(_dev(img), _dev(mask), radius).cpu().numpy()
        assert np.array_equal(got, io.inpaint_ns(img, mask, radius)), (n, radius)
