"""Which arithmetic is torch.norm(dim) / torch.sum(dim) on this device?  (development probe, run on the GPU box)

Candidates for |v| of a 3-vector along a strided axis, as the reference computes the plane spacing
(mpi_rendering.py:68-71, :104-107):  A  sqrt(fma(z,z, fma(y,y, x*x)))   B  sqrt((x*x + y*y) + z*z), squares rounded.
For the flow composite (torch.sum over S of weights * flow, :131): sequential, 4 interleaved chains, or ATen's blocked
order (4 rows x 4 chains when S >= 64).
"""
import torch

torch.manual_seed(0)
dev = "cuda"
n = 1 << 20
v = (torch.rand(1, 5, 3, n, device=dev) - 0.5) * torch.tensor([1e-2, 3e-3, 2e-2], device=dev).view(1, 1, 3, 1)
ref = torch.norm(v, dim=2, keepdim=True)
x, y, z = v[:, :, 0:1].double(), v[:, :, 1:2].double(), v[:, :, 2:3].double()
f32 = lambda t: t.float().double()
A = torch.sqrt(f32(z * z + f32(y * y + f32(x * x)))).float()
Bv = torch.sqrt(f32(f32(f32(x * x) + f32(y * y)) + f32(z * z))).float()
C = torch.sqrt(f32(f32(x * x) + f32(f32(y * y) + f32(z * z)))).float()
for name, c in (("A fma chain", A), ("B ((x2+y2)+z2)", Bv), ("C (x2+(y2+z2))", C)):
    print(f"norm vs {name}: mismatching {(c != ref).float().mean().item():.3e}")

for S in (16, 32, 64, 128):
    w = torch.rand(1, S, 1, n // 4, device=dev) * 0.05
    f = (torch.rand(1, S, 2, n // 4, device=dev) - 0.3) * 300
    ref = torch.sum(w * f, dim=1)
    p = (w * f).double()                       # exact sums of the rounded products below, rounded to fp32 after every add

    def chain(idx):
        acc = torch.zeros_like(p[:, 0])
        for s in idx:
            acc = f32(acc + p[:, s])
        return acc

    seq = chain(range(S)).float()
    four = [chain(range(i, S, 4)) for i in range(4)]
    inter = f32(f32(f32(four[0] + four[1]) + four[2]) + four[3]).float()
    if S >= 64:
        rows = []
        for ty in range(4):
            a = [chain(range(ty + 4 * i, S, 16)) for i in range(4)]
            rows.append(f32(f32(f32(a[0] + a[1]) + a[2]) + a[3]))
        blocked = f32(f32(rows[0] + rows[2]) + f32(rows[1] + rows[3])).float()
    else:
        blocked = inter
    exact = p.sum(1).float()
    for name, c in (("sequential", seq), ("4 chains s%4", inter), ("ATen blocked", blocked), ("exact", exact)):
        d = (c - ref).abs()
        print(f"S={S} sum vs {name}: mismatching {(c != ref).float().mean().item():.3e} max abs {d.max().item():.3e}")
    mine = f32(f32(four[0] + four[1]) + f32(four[2] + four[3])).float()
    print(f"S={S} kernel's order ((0+1)+(2+3)) vs torch: max abs {(mine - ref).abs().max().item():.3e}; |sum| max {ref.abs().max().item():.1f}")
