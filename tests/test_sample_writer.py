"""Asynchronous sample writer: the files equal what the reference's synchronous sequence
(utils/utils.py:284-286 + gen_3dphoto_dynamic_v2.py:120-122) writes for the same sample (tests/golden/sample_pair.npz)."""
import os

import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "sample_pair.npz")))


def _check_files(out_dir, name, gold):
    import write_flow
    flow = write_flow.readFlow(os.path.join(out_dir, "flows", f"{name}.flo"))
    assert np.array_equal(flow, gold["flow_mix"])
    raw = open(os.path.join(out_dir, "flows", f"{name}.flo"), "rb").read()
    assert raw[:4] == np.float32(202021.25).tobytes() and len(raw) == 12 + gold["flow_mix"].nbytes
    assert np.array_equal(cv2.imread(os.path.join(out_dir, "dst_images", f"{name}.png")), gold["inpainted"])
    assert np.array_equal(cv2.imread(os.path.join(out_dir, "src_images", f"{name}.png")), gold["src_np"])


def test_host_tensors_give_the_reference_files(gold, tmp_path):
    from sample_writer import SampleWriter
    with SampleWriter(str(tmp_path), workers=2, depth=2) as w:
        for k in range(5):                                    # more samples than slots: the ring is reused
            w.submit(f"demo_{k}", gold["flow_mix"], gold["frame_mix"], gold["fill_mask"], gold["src_np"])
    for k in range(5):
        _check_files(str(tmp_path), f"demo_{k}", gold)


def test_errors_surface(gold, tmp_path):
    from sample_writer import SampleWriter
    w = SampleWriter(str(tmp_path), workers=1, depth=1)
    w.submit("bad", gold["flow_mix"], gold["frame_mix"][:, :, :2], gold["fill_mask"], gold["src_np"])   # 2-channel image
    with pytest.raises(Exception):
        w.close()


@pytest.mark.gpu
def test_device_tensors_are_copied_on_a_side_stream(gold, tmp_path):
    from sample_writer import SampleWriter
    dev = lambda k: torch.from_numpy(gold[k]).cuda()
    with SampleWriter(str(tmp_path), workers=2, depth=2) as w:
        for k in range(6):
            flow, frame, fill = dev("flow_mix"), dev("frame_mix"), dev("fill_mask")
            w.submit(f"demo_{k}", flow, frame, fill, gold["src_np"])
            del flow, frame, fill                             # the writer keeps what it needs alive
            torch.empty(1 << 20, device="cuda").fill_(k)      # allocator churn while copies are in flight
    for k in range(6):
        _check_files(str(tmp_path), f"demo_{k}", gold)


@pytest.mark.gpu
def test_device_inpaint_gives_the_same_files(gold, tmp_path):
    """hole filling on the GPU (batches of 4, a partial batch at the end): the files still equal the reference's"""
    from sample_writer import SampleWriter
    dev = lambda k: torch.from_numpy(gold[k]).cuda()
    with SampleWriter(str(tmp_path), workers=2, depth=2, device_inpaint=True, inpaint_batch=4) as w:
        for k in range(10):
            flow, frame, fill = dev("flow_mix"), dev("frame_mix"), dev("fill_mask")
            w.submit(f"demo_{k}", flow, frame, fill, gold["src_np"])
            frame.fill_(0)                                    # the caller may reuse its buffers at once
            del flow, frame, fill
    for k in range(10):
        _check_files(str(tmp_path), f"demo_{k}", gold)
