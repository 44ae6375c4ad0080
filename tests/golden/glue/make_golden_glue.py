"""Generates tests/golden/glue/ref_glue.npz by running the reference's OWN Python (imported unmodified from
/root/reference through refsrc.load, CPU, this container) for the per-iteration glue and loss around the operator
(SURVEY.md rows a17 / N2 / N3):

* transform_to_frame + transformed_params2rendervar + transformed_params2depthplussilhouette
      R/utils/slam_helpers.py:124-139,196-304   (values and gradients, isotropic / anisotropic, with / without
      camera gradient)
* calc_ssim, l1_loss_v1                          R/utils/slam_external.py:66-97, R/utils/slam_helpers.py:6-7
* get_loss(mapping=True) and get_loss(tracking=True, use_sil_for_loss=True)
      R/scripts/splatam.py:214-347 -- the UNMODIFIED function, run with a stand-in `Renderer` that returns
      preset images, so the fixture pins the masks, the L1 / SSIM terms, the loss weights and dLoss/d(rendered
      images) -- exactly what the fused loss kernels (csrc/train_ops.cu) must reproduce.

The reference hard-codes `.cuda()` / device="cuda"; for this CPU run both are patched to stay on the CPU.  Nothing is
copied into the repo: the functions run where they lie.   Run:  python tests/golden/glue/make_golden_glue.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
import refsrc  # noqa: E402


def cpu_patches():
    torch.Tensor.cuda = lambda self, *a, **k: self
    for name in ("zeros", "zeros_like", "ones", "eye"):
        orig = getattr(torch, name)
        setattr(torch, name, (lambda o: lambda *a, **k: o(*a, **{kk: vv for kk, vv in k.items()
                                                            if not (kk == "device" and str(vv).startswith("cuda"))}))(orig))


def make_params(g, P, aniso):
    mk = lambda *s: torch.randn(*s, generator=g)
    return dict(means3D=mk(P, 3) * 2, rgb_colors=torch.rand(P, 3, generator=g), unnorm_rotations=mk(P, 4),
                logit_opacities=mk(P, 1), log_scales=mk(P, 3 if aniso else 1) * 0.3 - 3.0,
                cam_unnorm_rots=torch.tensor([[1.0, 0.02, -0.03, 0.01], [0.9, -0.1, 0.2, 0.05]]).T.reshape(1, 4, 2).contiguous(),
                cam_trans=torch.tensor([[0.1, -0.2, 0.05], [-0.3, 0.1, 0.2]]).T.reshape(1, 3, 2).contiguous())


class FakeCamera:     # stands in for diff_gaussian_rasterization while the reference modules are imported
    pass


def main():
    cpu_patches()
    import types
    pkg = types.ModuleType("fake_rasterizer_pkg")
    pkg.GaussianRasterizer = object
    pkg.GaussianRasterizationSettings = FakeCamera
    R = refsrc.load(pkg)
    assert R.root == "/root/reference", "fixtures are generated from the reference tree itself"
    H_, E_, S_ = R.slam_helpers, R.slam_external, R.splatam
    out = {}

    # ---- glue: values and gradients --------------------------------------------------------------------------
    w2c0 = torch.eye(4)
    w2c0[2, :] = torch.tensor([0.02, -0.01, 0.999, 0.3])
    w2c0[0, :] = torch.tensor([0.999, 0.01, -0.02, -0.1])
    out["glue_w2c0"] = w2c0.numpy()
    for aniso in (False, True):
        for camgrad in (False, True):
            tag = "glue_%s_%s_" % ("aniso" if aniso else "iso", "cam" if camgrad else "nocam")
            g = torch.Generator().manual_seed(20 + 2 * aniso + camgrad)
            P = 800
            base = make_params(g, P, aniso)
            weights = [torch.randn(P, n, generator=g) for n in (3, 4, 1, 3, 3)]
            p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
            t = 1
            tg = H_.transform_to_frame(p, t, gaussians_grad=True, camera_grad=camgrad)
            rgb = H_.transformed_params2rendervar(p, tg)
            dep = H_.transformed_params2depthplussilhouette(p, w2c0, tg)
            outs = [rgb["means3D"], rgb["rotations"], rgb["opacities"], rgb["scales"], dep["colors_precomp"]]
            assert torch.equal(rgb["colors_precomp"], p["rgb_colors"])
            # the depth rendervar shares means / rotations / opacities / scales with the RGB one: both raster calls
            # send gradient into them, so weight them twice (1.0 and 0.5)
            loss = sum((o * w).sum() for o, w in zip(outs, weights)) + 0.5 * sum(
                (o * w).sum() for o, w in zip([dep["means3D"], dep["rotations"], dep["opacities"], dep["scales"]], weights))
            loss.backward()
            for k, v in base.items():
                out[tag + "in_" + k] = v.numpy()
            for i, w in enumerate(weights):
                out[tag + "w%d" % i] = w.numpy()
            for name, o in zip(["means3D", "rotations", "opacities", "scales", "depth_colors"], outs):
                out[tag + "out_" + name] = o.detach().numpy()
            for k in ["means3D", "unnorm_rotations", "logit_opacities", "log_scales"] + (["cam_unnorm_rots", "cam_trans"] if camgrad else []):
                out[tag + "grad_" + k] = p[k].grad.numpy()
            out[tag + "time_idx"] = np.int64(t)

    # ---- SSIM / L1 -----------------------------------------------------------------------------------------------
    for i, shape in enumerate([(3, 45, 97), (3, 64, 80), (1, 11, 7)]):
        g = torch.Generator().manual_seed(40 + i)
        gt = torch.rand(*shape, generator=g)
        im = (gt + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1)
        a = im.clone().requires_grad_(True)
        s = E_.calc_ssim(a, gt)
        s.backward()
        b = im.clone().requires_grad_(True)
        l1 = H_.l1_loss_v1(b, gt)
        l1.backward()
        tag = "ssim%d_" % i
        out[tag + "im"], out[tag + "gt"] = im.numpy(), gt.numpy()
        out[tag + "ssim"], out[tag + "dssim"] = s.detach().numpy(), a.grad.numpy()
        out[tag + "l1"], out[tag + "dl1"] = l1.detach().numpy(), b.grad.numpy()

    # ---- the unmodified get_loss with preset renders ---------------------------------------------------------------
    g = torch.Generator().manual_seed(50)
    Hh, Ww, P = 45, 97, 64
    base = make_params(g, P, False)
    ds = torch.rand(3, Hh, Ww, generator=g)
    ds[0] = 0.5 + 2.0 * ds[0]
    ds[1] = 0.9 + 0.2 * torch.rand(Hh, Ww, generator=g)            # silhouette around the 0.99 threshold
    ds[2] = ds[0] ** 2 + 0.01 * torch.rand(Hh, Ww, generator=g)
    ds[0, 5, 7] = float("nan")
    gt_d = 0.5 + 2.0 * torch.rand(1, Hh, Ww, generator=g)
    gt_d[0, :3] = 0.0                                              # invalid-depth rows
    gt_d[0, 20, 30:40] = 40.0                                      # outliers (for ignore_outlier_depth_loss)
    gt_im = torch.rand(3, Hh, Ww, generator=g)
    im = (gt_im + 0.2 * torch.randn(3, Hh, Ww, generator=g)).clamp(0, 1)
    radius = torch.randint(0, 5, (P,), generator=g, dtype=torch.int32)
    for k, v in dict(im=im, depth_sil=ds, gt_im=gt_im, gt_depth=gt_d, radius=radius).items():
        out["loss_" + k] = v.numpy()
    for mode in ("mapping", "tracking", "tracking_nosil", "mapping_outlier"):
        ds_m = ds.clone()
        if mode == "mapping_outlier":      # a NaN depth makes the reference's median (and with it the whole loss) NaN
            ds_m[0, 5, 7] = 1.0
        leaves = dict(im=im.clone().requires_grad_(True), ds=ds_m.requires_grad_(True))

        class FakeRenderer:
            calls = []

            def __init__(self, raster_settings=None):
                pass

            def __call__(self, **rv):
                FakeRenderer.calls.append(sorted(rv.keys()))
                first = len(FakeRenderer.calls) == 1
                return (leaves["im"] if first else leaves["ds"]), radius, None
        S_.Renderer = FakeRenderer
        p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        variables = dict(max_2D_radius=torch.zeros(P), means2D_gradient_accum=torch.zeros(P), denom=torch.zeros(P))
        curr = dict(cam=None, im=gt_im, depth=gt_d, id=1, intrinsics=None, w2c=torch.eye(4), iter_gt_w2c_list=None)
        tracking = mode.startswith("tracking")
        loss, variables, wl = S_.get_loss(
            p, curr, variables, 1, dict(im=0.5, depth=1.0), use_sil_for_loss=(mode == "tracking"), sil_thres=0.99,
            use_l1=True, ignore_outlier_depth_loss=(mode == "mapping_outlier"), tracking=tracking, mapping=not tracking)
        loss.backward()
        assert FakeRenderer.calls[0] == sorted(["means3D", "colors_precomp", "rotations", "opacities", "scales", "means2D"])
        tag = "loss_%s_" % mode
        out[tag + "loss"] = loss.detach().numpy()
        assert np.isfinite(out[tag + "loss"])
        out[tag + "w_im"], out[tag + "w_depth"] = wl["im"].detach().numpy(), wl["depth"].detach().numpy()
        out[tag + "d_im"] = leaves["im"].grad.numpy()
        out[tag + "d_ds"] = torch.nan_to_num(leaves["ds"].grad).numpy()
        out[tag + "seen"] = variables["seen"].numpy()
        out[tag + "max_2D_radius"] = variables["max_2D_radius"].numpy()

    path = os.path.join(HERE, "ref_glue.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%d arrays, %.1f KB" % (len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
