import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
case = sys.argv[1:8]
for plane in (0, 7, 14, 15):
    outs = []
    for e in ("pk", "tma"):
        f = tempfile.mktemp(suffix=".npz")
        env = dict(os.environ, DBG_OUT=f, MPIFLOW_RENDER_PATH=e, MPIFLOW_PK_DBG=hex(0x08000000 | (plane << 16)))
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dbg_case.py"), *case], env=env, check=True, stdout=subprocess.DEVNULL)
        z = np.load(f); outs.append(z[z.files[0]])
    a, b = outs
    m = (np.abs(a[0, 0] - b[0, 0]) < 1e-2) & (a[0, 0] > 1.5)
    for c, n in ((0, "ix"), (1, "iy")):
        ne_ = (a[0, c].view(np.int32) != b[0, c].view(np.int32)) & m
        print(f"   {n} differs at {ne_.sum()} of {m.sum()} max {np.abs((a[0,c]-b[0,c])*m).max():.3e}")
    d = np.abs(a[0, 2] - b[0, 2]) * m
    ne = (a[0, 2].view(np.int32) != b[0, 2].view(np.int32)) & m
    print(f"plane {plane}: both-lean pixels {m.mean():.3f}; sg differs at {ne.sum()} of {m.sum()}, max diff {d.max():.3e}, median nonzero diff {np.median(d[ne]) if ne.any() else 0:.3e}")
    ys, xs = np.nonzero(ne)
    print("   x histogram (16-col bins):", np.bincount(xs // 16, minlength=20).tolist())
    print("   y histogram (4-row bins):", np.bincount(ys // 4, minlength=24).tolist())
    k = np.argmax(d); yy, xx = divmod(int(k), d.shape[1]); print(f"   worst at y{yy} x{xx}: pk {a[0,2,yy,xx]!r} tma {b[0,2,yy,xx]!r} ix {a[0,0,yy,xx]!r} {b[0,0,yy,xx]!r} iy {a[0,1,yy,xx]!r} {b[0,1,yy,xx]!r}")
    for c, n in enumerate([]):
        eq = a[0, c].view(np.int32) == b[0, c].view(np.int32)
        d = np.abs(a[0, c] - b[0, c])
        print(f"{n:12s} identical {eq.mean():.3f}  max diff {d.max():.3e} rel {np.nanmax(d / (np.abs(b[0, c]) + 1e-30)):.3e}")
