#!/usr/bin/env python
"""Small fwd+bwd runs (plain, fused two-set, sync-free) for compute-sanitizer memcheck / racecheck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import scenes  # noqa: E402
import splatam_b200 as S  # noqa: E402

dev = torch.device("cuda:0")
for make in (scenes.config1, scenes.edge_cases, lambda: scenes.dense_opaque(P=1500, w=96, h=64)):
    sc = make()
    rs = sc.settings(S.GaussianRasterizationSettings, dev)
    dL = torch.randn(3, sc.h, sc.w).to(dev)
    a = sc.inputs(dev, requires_grad=True)
    c, r, d = S.GaussianRasterizer(rs)(**a)
    c.backward(dL)
    R = c.grad_fn.state.num_rendered
    b = sc.inputs(dev, requires_grad=True)
    ex = torch.rand(sc.P, 3, device=dev, requires_grad=True)
    f1, f2, _, _ = S.GaussianRasterizer(rs).forward_fused(means3D=b["means3D"], means2D=b["means2D"], opacities=b["opacities"],
                                                          colors_precomp=b["colors_precomp"], colors_extra=ex,
                                                          scales=b["scales"], rotations=b["rotations"])
    (f1.sum() + f2.sum()).backward()
    e = sc.inputs(dev, requires_grad=True)
    c2, _, _ = S.GaussianRasterizer(rs, max_rendered=R + 100)(**e)
    c2.backward(dL)
    torch.cuda.synchronize()
    print(sc.name, "R", R, "ok", bool(torch.equal(c, c2)))
