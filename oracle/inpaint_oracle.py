"""Python front end of oracle/inpaint_oracle.c (cv2.inpaint(INPAINT_NS) restated) -- TEST INFRASTRUCTURE ONLY;
nothing under mpi-flow_b200/ imports it.  Reference call site: utils/utils.py:284-286."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libinpaint_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make -C oracle all`")
        _lib = ctypes.CDLL(path)
        _lib.orc_inpaint_ns.restype = ctypes.c_int
        _lib.orc_inpaint_ns.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p]
    return _lib


def inpaint_ns(img, mask, radius=3, variant=0):
    """img HxWx3 uint8, mask HxW uint8 (non-zero = hole) -> HxWx3 uint8, what cv2.inpaint(img, mask, radius, INPAINT_NS)
    returns.  variant != 0: other readings of the border / arithmetic details (see the C file)."""
    img = np.ascontiguousarray(img, np.uint8)
    mask = np.ascontiguousarray(mask, np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3 and mask.shape == img.shape[:2]
    out = np.empty_like(img)
    rc = _load().orc_inpaint_ns(img.ctypes.data, mask.ctypes.data, img.shape[0], img.shape[1], int(round(radius)), variant,
                                out.ctypes.data)
    assert rc == 0
    return out
