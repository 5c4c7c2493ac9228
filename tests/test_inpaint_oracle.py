"""oracle/inpaint_oracle.c (cv2.inpaint(INPAINT_NS) restated; SURVEY.md 8(f) row 4, reference utils/utils.py:284-286)
pinned against cv2's own outputs: the committed fixtures tests/golden/inpaint.npz (made by
tests/golden/make_inpaint_golden.py), the `inpainted` image the unmodified reference produced for the sample fixture,
and -- when cv2 is importable -- cv2.inpaint itself on fresh random cases."""
import os

import numpy as np
import pytest

from oracle import inpaint_oracle as io


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "inpaint.npz"))


def test_oracle_equals_cv2_fixtures_bit_for_bit(gold):
    n = int(gold["count"])
    assert n >= 30
    filled = 0
    for i in range(n):
        got = io.inpaint_ns(gold[f"img{i}"], gold[f"mask{i}"], int(gold[f"radius{i}"]))
        assert np.array_equal(got, gold[f"out{i}"]), f"case {i}"
        filled += int((gold[f"mask{i}"] > 0).sum())
    assert filled > 40000           # the fixtures are not trivially small


def test_fixtures_discriminate_the_other_readings(gold):
    """each alternative reading of the border / arithmetic details differs from cv2 somewhere in the fixtures"""
    n = int(gold["count"])
    for variant in (1, 2, 4):
        diff = sum(int((io.inpaint_ns(gold[f"img{i}"], gold[f"mask{i}"], int(gold[f"radius{i}"]), variant)
                        != gold[f"out{i}"]).sum()) for i in range(n))
        assert diff > 0, f"variant {variant} is not told apart"


def test_oracle_reproduces_the_reference_sample(golden_dir):
    """`inpainted` of sample_pair.npz is what the unmodified reference's cv2.inpaint call returned"""
    g = np.load(os.path.join(golden_dir, "sample_pair.npz"))
    assert np.array_equal(io.inpaint_ns(g["frame_mix"], g["fill_mask"], 3), g["inpainted"])


def test_nothing_to_fill_and_everything_to_fill():
    rng = np.random.default_rng(5)
    img = (rng.random((20, 30, 3)) * 255).astype(np.uint8)
    assert np.array_equal(io.inpaint_ns(img, np.zeros((20, 30), np.uint8)), img)
    # no known pixel at all: the band is empty, nothing is ever popped, the image stays as it is (cv2 does the same)
    assert np.array_equal(io.inpaint_ns(img, np.ones((20, 30), np.uint8)), img)


def test_oracle_equals_live_cv2_on_random_cases():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(77)
    total = 0
    for trial in range(40):
        H, W = int(rng.integers(24, 120)), int(rng.integers(24, 160))
        img = (rng.random((H, W, 3)) * 255).astype(np.uint8)
        if trial % 3:
            img = cv2.GaussianBlur(img, (0, 0), float(rng.uniform(0.8, 4)))
        mask = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(2, 20))):
            x, y = int(rng.integers(0, W)), int(rng.integers(0, H))
            cv2.line(mask, (x, y), (int(np.clip(x + rng.integers(-30, 30), 0, W - 1)),
                                    int(np.clip(y + rng.integers(-30, 30), 0, H - 1))), 255, int(rng.integers(1, 7)))
        if trial % 5 == 0:
            mask = ((rng.random((H, W)) < 0.3) * 255).astype(np.uint8)
        radius = 3 if trial % 4 else int(rng.integers(1, 6))
        assert np.array_equal(io.inpaint_ns(img, mask, radius), cv2.inpaint(img, mask, radius, cv2.INPAINT_NS)), trial
        total += int((mask > 0).sum())
    assert total > 20000
