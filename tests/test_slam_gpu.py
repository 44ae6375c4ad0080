"""BASELINE configs 4/5 stand-in (the datasets are not available offline): a slim SplaTAM-style SLAM run --
tracking every frame, keyframe mapping -- over a synthetic RGB-D sequence, with this repo's fused path, its plain
operator path, and the unmodified reference extension.  End metrics (ATE-RMSE, PSNR) must agree."""
import numpy as np
import pytest
import torch

import scenes
from util import reference_extension

pytestmark = pytest.mark.gpu


def _sequence(dev, Settings, T=9):
    from splatam_b200 import slam
    sc = scenes.room(seed=31, P=30_000, cam=dict(w=320, h=192, fx=160.0, fy=160.0, cx=159.5, cy=95.5))
    cam = sc.settings(Settings, dev)
    gt = dict(means3D=sc.means3D, rgb_colors=sc.colors, unnorm_rotations=sc.rotations,
              logit_opacities=torch.logit(sc.opacities.clamp(0.02, 0.98)), log_scales=torch.log(sc.scales[:, :1]))
    gt = {k: v.to(dev) for k, v in gt.items()}
    rots, trans = slam.look_trajectory(T, dev)
    g = torch.Generator().manual_seed(5)
    init = dict(gt)
    init["rgb_colors"] = (gt["rgb_colors"] + 0.08 * torch.randn(gt["rgb_colors"].shape, generator=g).to(dev)).clamp(0, 1)
    init["log_scales"] = gt["log_scales"] + 0.05 * torch.randn(gt["log_scales"].shape, generator=g).to(dev)
    return sc, cam, gt, init, rots, trans


def test_slim_slam_ate_and_psnr_match_reference(cuda_device):
    import splatam_b200 as S
    from splatam_b200 import slam
    dev = cuda_device
    sc, cam, gt, init, rots_gt, trans_gt = _sequence(dev, S.GaussianRasterizationSettings)
    frames = [slam.render_frame(gt, rots_gt, trans_gt, t, cam) for t in range(rots_gt.shape[-1])]
    out = {}
    out["fused"] = slam.run_slam(init, frames, cam, fused=True)
    out["plain"] = slam.run_slam(init, frames, cam, fused=False)
    ref = reference_extension()
    if ref is not None:
        rcam = sc.settings(ref.GaussianRasterizationSettings, dev)
        rframes = [dict(f, cam=rcam) for f in frames]
        render = lambda settings, **rv: ref.GaussianRasterizer(raster_settings=settings)(**rv)
        out["reference"] = slam.run_slam(init, rframes, rcam, render=render, fused=False)
    ate = {k: slam.ate_rmse(v["rots"], v["trans"], rots_gt, trans_gt) for k, v in out.items()}
    ps = {k: v["psnr"] for k, v in out.items()}
    print("ATE-RMSE [m]", ate, "PSNR [dB]", ps)
    travelled = float(trans_gt[0, :, -1].norm())
    for k in out:
        assert ate[k] < 0.03 * travelled, (k, ate[k], travelled)       # tracks within 3 % of the path length
        assert ps[k] > 28.0, (k, ps[k])
    # run-to-run spread of ONE implementation is already ~1.3 dB (the reference arm alone gave 32.8 and 34.1 dB on
    # two runs: float-atomic ordering feeds a 160-step optimisation), so the arms are compared at 2 dB
    base = "reference" if "reference" in out else "plain"
    for k in out:
        assert abs(ps[k] - ps[base]) < 2.0, (k, ps)
        assert abs(ate[k] - ate[base]) < 0.01 * travelled, (k, ate)
