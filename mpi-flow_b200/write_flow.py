"""Flow / depth file formats -- drop-in for MPI-Flow's write_flow.py (the wire format the
generator emits, gen_3dphoto_dynamic_v2.py:120-122).  Written fresh; same function names, arguments and bytes.

  .flo (Middlebury)   float32 tag 202021.25, int32 width, int32 height, then rows of interleaved (u, v) float32
  KITTI flow PNG      uint16 BGR image: (valid, 64 v + 2^15, 64 u + 2^15)
  .pfm                'PF' / 'Pf' header, dims, scale (sign = endianness), rows bottom-up
Little-endian hosts only, like the reference (write_flow.py:19).
"""
import re
from os.path import splitext

import numpy as np

TAG_FLOAT = 202021.25
TAG_CHAR = np.array([TAG_FLOAT], np.float32)


def writeFlow(filename, uv, v=None):
    """uv [h,w,2] (or u and v separately, [h,w] each) -> Middlebury .flo (reference :74-103)"""
    if v is None:
        assert uv.ndim == 3 and uv.shape[2] == 2
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u = uv
    assert u.shape == v.shape
    height, width = u.shape
    with open(filename, "wb") as f:
        f.write(TAG_CHAR.tobytes())
        f.write(np.array([width, height], np.int32).tobytes())
        f.write(np.stack([u, v], axis=-1).astype(np.float32).tobytes())   # u, v interleaved along the row


def readFlow(fn):
    """Middlebury .flo -> [h,w,2] float32, None (and a message) on a wrong tag (reference :14-33)"""
    with open(fn, "rb") as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != np.float32(TAG_FLOAT):
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.fromfile(f, np.int32, count=1)[0])
        h = int(np.fromfile(f, np.int32, count=1)[0])
        data = np.fromfile(f, np.float32, count=2 * w * h)
        return np.resize(data, (h, w, 2))


def readPFM(file):
    """.pfm -> [h,w,3] or [h,w] float32, top row first (reference :36-71)"""
    with open(file, "rb") as f:
        header = f.readline().rstrip()
        if header not in (b"PF", b"Pf"):
            raise Exception("Not a PFM file.")
        color = header == b"PF"
        m = re.match(rb"^(\d+)\s(\d+)\s$", f.readline())
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().rstrip())
        data = np.fromfile(f, ("<" if scale < 0 else ">") + "f")
    return np.flipud(np.reshape(data, (height, width, 3) if color else (height, width)))


def readFlowKITTI(filename):
    """KITTI 16-bit flow PNG -> (flow [h,w,2] float32, valid [h,w] float32) (reference :106-111)"""
    import cv2
    img = cv2.imread(filename, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32)
    return (img[:, :, :2] - 2 ** 15) / 64.0, img[:, :, 2]


def readDispKITTI(filename):
    """KITTI 16-bit disparity PNG -> (flow (-disp, 0) [h,w,2], valid [h,w] bool) (reference :114-118)"""
    import cv2
    disp = cv2.imread(filename, cv2.IMREAD_ANYDEPTH) / 256.0
    return np.stack([-disp, np.zeros_like(disp)], -1), disp > 0.0


def writeFlowKITTI(filename, uv):
    """flow [h,w,2] -> KITTI 16-bit PNG with valid = 1 everywhere (reference :121-125)"""
    import cv2
    enc = np.concatenate([64.0 * uv + 2 ** 15, np.ones([uv.shape[0], uv.shape[1], 1])], axis=-1).astype(np.uint16)
    cv2.imwrite(filename, enc[..., ::-1])


def read_gen(file_name, pil=False):
    """dispatch on the extension (reference :128-143)"""
    ext = splitext(file_name)[-1]
    if ext in (".png", ".jpeg", ".ppm", ".jpg"):
        from PIL import Image
        return Image.open(file_name)
    if ext in (".bin", ".raw"):
        return np.load(file_name)
    if ext == ".flo":
        return readFlow(file_name).astype(np.float32)
    if ext == ".pfm":
        flow = readPFM(file_name).astype(np.float32)
        return flow if flow.ndim == 2 else flow[:, :, :-1]
    raise ValueError("wrong file type: %s" % ext)


def depth_read(filename):
    """Sintel-style depth file (.dpt: tag, width, height, float32 rows) -> [h,w] (reference :146-157)"""
    with open(filename, "rb") as f:
        check = np.fromfile(f, dtype=np.float32, count=1)[0]
        assert check == np.float32(TAG_FLOAT), f" depth_read:: Wrong tag in flow file (should be: {TAG_FLOAT}, is: {check}). Big-endian machine? "
        width = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        height = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        size = width * height
        assert width > 0 and height > 0 and 1 < size < 100000000, f" depth_read:: Wrong input size (width = {width}, height = {height})."
        return np.fromfile(f, dtype=np.float32, count=-1).reshape((height, width))
