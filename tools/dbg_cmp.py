"""Debug aid: the same case rendered in two subprocesses with different MPIFLOW_* settings; lists the pixels
where the outputs differ.  usage: dbg_cmp.py "ENV_A" "ENV_B" B S H W kind pose seed [thr]"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ea, eb = sys.argv[1], sys.argv[2]
case = sys.argv[3:10]
thr = float(sys.argv[10]) if len(sys.argv) > 10 else 1e-5
outs = []
for e in (ea, eb):
    f = tempfile.mktemp(suffix=".npz")
    env = dict(os.environ, DBG_OUT=f)
    for kv in e.split():
        k, v = kv.split("=", 1); env[k] = v
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dbg_case.py"), *case], env=env, check=True, stdout=subprocess.DEVNULL)
    z = np.load(f); outs.append([z[k] for k in z.files])
for n, a, b in zip(["rgb", "depth", "mask", "flow", "obj"], *outs):
    d = np.abs(a - b) / (1 + np.abs(b) if n == "depth" else 1)
    dm = d.max(axis=1)  # over channels
    bad = np.argwhere(dm > thr)
    print(f"{n:6s} max diff {d.max():.3e}  #pixels > {thr:g}: {len(bad)} of {dm.size}")
    for bb, y, x in bad[:24]:
        print(f"   b{bb} y{y} x{x}  diff {dm[bb, y, x]:.3e}   (tile row {y % 16}, lane {x % 32})")
a, b = outs
eq = (a[0] == b[0]).all(axis=1)[0]
print("rgb bit-identical pixels: %.3f" % eq.mean())
H, W = eq.shape
for y in range(0, H, 4):
    print("%3d " % y + "".join("." if eq[y:y + 4, x:x + 8].all() else ("x" if not eq[y:y + 4, x:x + 8].any() else "+") for x in range(0, W, 8)))
