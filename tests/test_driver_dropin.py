"""The drop-in boundary of north_star: `gen_3dphoto_dynamic_v2.py` must run UNCHANGED on the package.

* CPU: the generator's import block (gen_3dphoto_dynamic_v2.py:8-14) resolves against mpi-flow_b200/ first on
  sys.path; the two loaders equal the reference's (torchvision ToTensor, cv2.imread/255, utils/utils.py:35-52);
  `utils` stays a namespace package, so the reference's other `utils.*` modules are not shadowed.
* GPU: the unmodified script (byte copy in oracle/_ref/driver/, made by oracle/Makefile) is executed in a scratch
  directory with a stub `model.AdaMPI.MPIPredictor` (the CNN is out of scope) in its production dtype -- every
  tensor it hands to render_3dphoto_dynamic is .half() (gen_3dphoto_dynamic_v2.py:46,59,82-84,104) -- and must
  write the three files per sample (:120-122).
"""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mpi-flow_b200")
DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver", "gen_3dphoto_dynamic_v2.py")

STUB_MODEL = '''
"""Stand-in for the AdaMPI network (out of scope): a deterministic MPI from the image and the disparity."""
import torch


class MPIPredictor(torch.nn.Module):
    def __init__(self, width, height, num_planes):
        super().__init__()
        self.num_planes = num_planes
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, image, disp):
        b, _, h, w = image.shape
        S = self.num_planes
        d = torch.linspace(1, 0.001, S + 2, device=image.device)[1:-1].to(image.dtype)       # model/AdaMPI.py:64-68
        sigma = torch.relu(1 - (disp.unsqueeze(1) - d.view(1, S, 1, 1, 1)).abs() * S) * 4 + 1e-4
        rgb = image.unsqueeze(1).expand(b, S, 3, h, w) * (0.5 + 0.5 * d.view(1, S, 1, 1, 1))
        return torch.cat((rgb, sigma), 2).to(image.dtype), d.unsqueeze(0).expand(b, S).contiguous()
'''


def _write_inputs(base, h=96, w=160):
    import cv2
    rng = np.random.default_rng(7)
    for sub in ("images", "disps", "masks"):
        os.makedirs(os.path.join(base, sub))
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // w), (yy * 255 // h), ((xx + yy) % 256)], -1).astype(np.uint8)
    img = cv2.GaussianBlur(img + rng.integers(0, 20, img.shape).astype(np.uint8), (5, 5), 0)
    disp = (40 + 180 * (yy / h)).astype(np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[20:60, 30:80] = 1
    mask[50:90, 100:150] = 2
    for name in ("a.png", "b.png"):
        cv2.imwrite(os.path.join(base, "images", name), img)
        cv2.imwrite(os.path.join(base, "disps", name), disp)
        cv2.imwrite(os.path.join(base, "masks", name), mask)
        img = img[:, ::-1].copy()


def test_generator_import_block_resolves_on_the_package(tmp_path):
    """gen_3dphoto_dynamic_v2.py:8-14, executed in a fresh interpreter with the package first on the path and a
    directory holding another `utils.*` module after it (the reference's utils/ has no __init__.py either)."""
    other = tmp_path / "ref_like" / "utils"
    other.mkdir(parents=True)
    (other / "transform.py").write_text("MARK = 'reference-side module'\n")
    code = textwrap.dedent("""
        from write_flow import writeFlow
        from utils.utils import (
            image_to_tensor,
            disparity_to_tensor,
            render_3dphoto_dynamic,
        )
        import utils.utils, utils.transform, os
        assert os.path.dirname(utils.utils.__file__).endswith(os.path.join("mpi-flow_b200", "utils"))
        assert utils.transform.MARK == 'reference-side module'
        from utils.mpi import mpi_rendering
        from utils.mpi.homography_sampler import HomographySample
        from utils.utils import arrowon, gen_swing_path, render_novel_view, generate_random_pose, render_novel_view_dynamic
        print("ok")
    """)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, str(tmp_path / "ref_like")]))
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=env, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_loaders_equal_the_reference_definitions(tmp_path):
    """utils/utils.py:35-52: ToTensor()(PIL RGB) and cv2.imread(path, 0) / 255 -> float."""
    import cv2
    import torch
    from PIL import Image
    from torchvision import transforms
    from utils import utils as U
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (17, 23, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "i.png"), img)
    cv2.imwrite(str(tmp_path / "d.png"), img[..., 0])
    want = transforms.ToTensor()(Image.open(str(tmp_path / "i.png")).convert("RGB"))
    got = U.image_to_tensor(str(tmp_path / "i.png"))
    assert got.dtype == torch.float32 and got.shape == (1, 3, 17, 23) and torch.equal(got[0], want)
    assert torch.equal(U.image_to_tensor(str(tmp_path / "i.png"), unsqueeze=False), want)
    dwant = torch.from_numpy(cv2.imread(str(tmp_path / "d.png"), 0) / 255)[None, ...].unsqueeze(0).float()
    dgot = U.disparity_to_tensor(str(tmp_path / "d.png"))
    assert dgot.dtype == torch.float32 and torch.equal(dgot, dwant)
    assert U.disparity_to_tensor(str(tmp_path / "d.png"), unsqueeze=False).shape == (1, 17, 23)
    flow = np.zeros((64, 64, 2), np.float32)
    flow[..., 0] = 5
    over = U.arrowon(np.zeros((64, 64, 3), np.uint8), flow)
    assert over.shape == (64, 64, 3) and over[..., 2].any() and not over[..., 0].any()


@pytest.mark.gpu
def test_unmodified_generator_runs_in_half_precision(tmp_path):
    import torch
    if not os.path.isfile(DRIVER):
        pytest.skip("oracle/_ref/driver is built by `make -C oracle ref` in the build container")
    stub = tmp_path / "stub"
    (stub / "model").mkdir(parents=True)
    (stub / "model" / "AdaMPI.py").write_text(STUB_MODEL)
    _write_inputs(str(tmp_path / "base"))
    torch.save({"num_planes": 8, "weight": {"dummy": torch.zeros(1)}}, str(tmp_path / "ck.pth"))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([PKG, str(stub)]))
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, DRIVER, "--width", "160", "--height", "96", "--repeat", "2", "--base",
                        str(tmp_path / "base"), "--out", str(out), "--ckpt_path", str(tmp_path / "ck.pth")],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    import cv2
    from write_flow import readFlow
    src_in = cv2.imread(str(tmp_path / "base" / "images" / "a.png"))
    for name in ("a", "b"):
        for rep in range(2):
            flo = readFlow(str(out / "flows" / f"{name}_{rep}.flo"))
            dst = cv2.imread(str(out / "dst_images" / f"{name}_{rep}.png"))
            src = cv2.imread(str(out / "src_images" / f"{name}_{rep}.png"))
            assert flo.shape == (96, 160, 2) and np.isfinite(flo).all() and np.abs(flo).max() > 0.5
            assert np.abs(flo).max() <= 200
            assert dst.shape == (96, 160, 3) and src.shape == (96, 160, 3)
            # the novel view is a small camera motion away from the photograph
            assert np.abs(dst.astype(np.int32) - src.astype(np.int32)).mean() < 40
        if name == "a":
            assert np.abs(src.astype(np.int32) - src_in.astype(np.int32)).max() <= 1    # fp16 round trip of the image
    # the same run with the hole filling on the GPU (MPIFLOW_INPAINT=device): the same files, byte for byte after decoding
    out2 = tmp_path / "out_device"
    r = subprocess.run([sys.executable, DRIVER, "--width", "160", "--height", "96", "--repeat", "2", "--base",
                        str(tmp_path / "base"), "--out", str(out2), "--ckpt_path", str(tmp_path / "ck.pth")],
                       cwd=str(tmp_path), env=dict(env, MPIFLOW_INPAINT="device"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    for name in ("a", "b"):
        for rep in range(2):
            for sub in ("dst_images", "src_images"):
                assert np.array_equal(cv2.imread(str(out / sub / f"{name}_{rep}.png")), cv2.imread(str(out2 / sub / f"{name}_{rep}.png")))
            assert np.array_equal(readFlow(str(out / "flows" / f"{name}_{rep}.flo")), readFlow(str(out2 / "flows" / f"{name}_{rep}.flo")))


def test_reference_modules_run_on_cpu_tensors_even_where_cuda_exists(monkeypatch):
    """bench.py --impl reference times the reference's CPU path ON THE GPU BOX: its inverse() hard-codes .cuda()
    (homography_sampler.py:15-17), so the shims have to follow the tensors' device, not the machine's."""
    import torch
    import synth
    from oracle import ref_modules, torch_ref
    if not ref_modules.available():
        pytest.skip("oracle/_ref/refsrc is built by `make -C oracle ref` in the build container")
    inp = {k: torch.from_numpy(v) for k, v in synth.frame_inputs(1, 4, 24, 32, "smooth", "random", 3).items()}
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)

    def no_cuda(self, *a, **k):
        raise AssertionError("a CPU run of the reference must not move tensors to CUDA")
    monkeypatch.setattr(torch.Tensor, "cuda", no_cuda)
    out = ref_modules.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"])
    want = torch_ref.render_frame(inp["mpi"], inp["disparity"], inp["G"], inp["Kinv"], inp["K"], inp["obj_mask"])
    for a, b in zip(out, want):
        assert torch.equal(a, b)
    assert torch.Tensor.cuda is no_cuda          # restored after the call
