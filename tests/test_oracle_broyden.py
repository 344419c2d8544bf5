"""Row f4: the optional Broyden root finder.  The reference has none, so there is no golden vector
to pin against — these are the defining properties of the CPU statement (oracle/port.py:deform_broyden) that the CUDA
path (mp_deform_broyden) is compared with in tests/test_gpu_parity.py."""
import torch

from multiply_b200 import scene as S
from oracle import port


def _points(p, n=3000, seed=0, sigma=0.03):
    g = torch.Generator().manual_seed(seed)
    v = p["verts_p"]
    return v[torch.randint(0, v.shape[0], (n,), generator=g)] + sigma * torch.randn(n, 3, generator=g)


def test_broyden_properties():
    sc = S.make_scene(P=2, S=16, seed=42)
    p = sc["persons"][1]
    x = _points(p)
    xc0, outl0 = port.deform_inverse(x, p)
    r0 = (port.forward_skinning(xc0, p)[0] - x).norm(dim=-1)
    xc, res, conv, outl = port.deform_broyden(x, p, max_steps=10, cvg_threshold=1e-5)
    assert torch.equal(outl, outl0)
    # the residual reported is the residual of the point returned
    assert torch.equal((port.forward_skinning(xc, p)[0] - x).norm(dim=-1), res)
    # never worse than the closed form; consistent points are left exactly where they were
    assert bool((res <= r0).all())
    same = r0 < 1e-5
    assert 0.3 < float(same.float().mean()) < 0.95          # the scene has both kinds of points
    assert torch.equal(xc[same], xc0[same])
    # most of the inconsistent points have a root nearby and reach it
    assert float(conv[~same].float().mean()) > 0.5
    assert float(conv.float().mean()) > 0.85
    assert bool((res[conv] < 1e-5).all())
    # zero steps = the closed-form inverse
    xz, rz, _, _ = port.deform_broyden(x, p, max_steps=0)
    assert torch.equal(xz, xc0) and torch.equal(rz, r0)


def test_root_finder_switch_in_deform_inverse():
    sc = S.make_scene(P=2, S=16, seed=42)
    p = sc["persons"][0]
    x = _points(p, n=1000, seed=3, sigma=0.06)
    xc0, outl = port.deform_inverse(x, p)
    xr, _, _, _ = port.deform_broyden(x, p, 10, 1e-5)
    xc1, outl1 = port.deform_inverse(x, dict(p, root_finder=(10, 1e-5)))
    assert torch.equal(outl, outl1) and bool(outl.any()) and bool((~outl).any())
    assert torch.equal(xc1[outl], xc0[outl])                 # outliers keep the closed form
    assert torch.equal(xc1[~outl], xr[~outl])
