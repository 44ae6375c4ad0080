"""GPU numerics of the fused training ops against plain PyTorch fp32 references of the same ops."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam(cuda_device):
    from splatam_b200.train_ops import FusedAdam
    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    sizes, lrs = [3000, 3000, 4001, 999, 1000], [1e-4, 2.5e-3, 1e-3, 5e-2, 1e-3]
    n = sum(sizes)
    init = torch.randn(n, generator=g)
    flat, grad = init.clone().to(dev), torch.zeros(n, device=dev)
    fused = FusedAdam(flat, grad, sizes, lrs, eps=1e-15)
    refs = [t.clone().to(dev).requires_grad_(True) for t in torch.split(init, sizes)]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(refs, lrs)], lr=0.0, eps=1e-15)
    for it in range(25):
        gr = torch.randn(n, generator=g) * (10.0 ** torch.randint(-4, 2, (1,), generator=g).item())
        grad.copy_(gr.to(dev))
        for p, gp in zip(refs, torch.split(gr.to(dev), sizes)):
            p.grad = gp.clone()
        fused.step()
        ref.step()
    out = torch.cat([p.detach() for p in refs])
    assert torch.allclose(flat, out, rtol=2e-5, atol=1e-7), float((flat - out).abs().max())
    # the kernel follows torch's op order with the same double->float scalar rounding: moments agree to ~1 ulp
    assert torch.allclose(fused.m, torch.cat([ref.state[p]["exp_avg"] for p in refs]), rtol=5e-6, atol=1e-9)
    assert torch.allclose(fused.v, torch.cat([ref.state[p]["exp_avg_sq"] for p in refs]), rtol=5e-6, atol=1e-12)


@pytest.mark.parametrize("shape", [(3, 680, 1200), (3, 45, 97), (1, 11, 7)])
def test_fused_image_loss_matches_torch(shape, cuda_device):
    from splatam_b200.mapping import calc_ssim
    from splatam_b200.train_ops import image_loss
    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    gt = torch.rand(*shape, generator=g).to(dev)
    im = (gt.cpu() + 0.2 * torch.randn(*shape, generator=g)).clamp(0, 1).to(dev)
    a = im.clone().requires_grad_(True)
    la = image_loss(a, gt)
    (la * 3.0).backward()
    b = im.clone().requires_grad_(True)
    lb = 0.8 * torch.abs(b - gt).mean() + 0.2 * (1.0 - calc_ssim(b, gt))
    (lb * 3.0).backward()
    assert abs(float(la.detach()) - float(lb.detach())) < 2e-6 * max(1.0, abs(float(lb))), (float(la), float(lb))
    err = (a.grad - b.grad).norm() / b.grad.norm()
    assert err < 2e-4, float(err)
    # the SSIM part alone (L1's sign() dominates the gradient norm otherwise)
    a2 = im.clone().requires_grad_(True); image_loss(a2, gt, 0.0, 1.0).backward()
    b2 = im.clone().requires_grad_(True); (1.0 - calc_ssim(b2, gt)).backward()
    err2 = (a2.grad - b2.grad).norm() / b2.grad.norm()
    assert err2 < 5e-4, float(err2)


@pytest.mark.parametrize("aniso,camgrad", [(False, False), (True, True), (False, True)])
def test_prepare_gaussians_matches_torch_glue(aniso, camgrad, cuda_device):
    """Fused glue kernel vs the PyTorch restatement of R/utils/slam_helpers.py (values and all gradients)."""
    from splatam_b200 import mapping as M
    dev = cuda_device
    g = torch.Generator().manual_seed(2)
    P = 5000
    mk = lambda *s: torch.randn(*s, generator=g).to(dev)
    base = dict(means3D=mk(P, 3) * 2, rgb_colors=torch.rand(P, 3, generator=g).to(dev), unnorm_rotations=mk(P, 4),
                logit_opacities=mk(P, 1), log_scales=mk(P, 3 if aniso else 1) * 0.3 - 3.0,
                cam_unnorm_rots=torch.tensor([[1.0, 0.02, -0.03, 0.01]]).T.reshape(1, 4, 1).to(dev),
                cam_trans=torch.tensor([0.1, -0.2, 0.05]).reshape(1, 3, 1).to(dev))
    w2c0 = torch.eye(4, device=dev)
    w2c0[2, :] = torch.tensor([0.02, -0.01, 0.999, 0.3], device=dev)
    weights = [mk(P, 3), mk(P, 4), mk(P, 1), mk(P, 3), mk(P, 3)]

    def run(fused):
        p = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        if fused:
            rgb, dep = M.fused_rendervars(p, 0, w2c0, camera_grad=camgrad)
        else:
            tg = M.transform_to_frame(p, 0, gaussians_grad=True, camera_grad=camgrad)
            rgb, dep = M.rgb_rendervar(p, tg), M.depth_sil_rendervar(p, w2c0, tg)
        outs = [rgb["means3D"], rgb["rotations"], rgb["opacities"], rgb["scales"], dep["colors_precomp"]]
        # the depth rendervar shares means/rot/opac/scales with the rgb one: weight them twice like two raster calls
        loss = sum((o * w).sum() for o, w in zip(outs, weights)) + 0.5 * sum(
            (o * w).sum() for o, w in zip([dep["means3D"], dep["rotations"], dep["opacities"], dep["scales"]], weights))
        loss.backward()
        return outs, p

    (fo, fp), (to, tp) = run(True), run(False)
    for a, b in zip(fo, to):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
    keys = ["means3D", "unnorm_rotations", "logit_opacities", "log_scales"] + (["cam_unnorm_rots", "cam_trans"] if camgrad else [])
    for k in keys:
        a, b = fp[k].grad, tp[k].grad
        assert a is not None and b is not None, k
        err = (a - b).norm() / b.norm().clamp_min(1e-20)
        assert err < 1e-4, (k, float(err))


@pytest.mark.parametrize("use_sil,depth_mean,with_im", [(False, True, False), (True, False, True)])
def test_masked_l1_matches_torch(use_sil, depth_mean, with_im, cuda_device):
    from splatam_b200.train_ops import masked_l1
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    H, W = 67, 131
    ds = torch.rand(3, H, W, generator=g)
    ds[1] = 0.9 + 0.2 * torch.rand(H, W, generator=g)         # silhouette around the 0.99 threshold
    ds[0, 5, 7] = float("nan")
    gt_d = 2.0 * torch.rand(1, H, W, generator=g)
    gt_d[0, :3] = 0.0                                         # invalid depth rows
    im, gt_im = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    ds, gt_d, im, gt_im = [t.to(dev) for t in (ds, gt_d, im, gt_im)]

    a_ds, a_im = ds.clone().requires_grad_(True), im.clone().requires_grad_(True)
    ld, li = masked_l1(a_ds, gt_d, a_im if with_im else None, gt_im if with_im else None, sil_thres=0.99,
                       use_sil=use_sil, depth_mean=depth_mean)
    (2.0 * ld + (0.5 * li if with_im else 0.0)).backward()

    b_ds, b_im = ds.clone().requires_grad_(True), im.clone().requires_grad_(True)
    depth = b_ds[0:1]
    unc = (b_ds[2:3] - depth ** 2).detach()
    mask = (gt_d > 0) & (~torch.isnan(depth)) & (~torch.isnan(unc))
    if use_sil:
        mask = mask & (b_ds[1] > 0.99)
    mask = mask.detach()
    rd = torch.abs(gt_d - depth)[mask].mean() if depth_mean else torch.abs(gt_d - depth)[mask].sum()
    ri = torch.abs(gt_im - b_im)[torch.tile(mask, (3, 1, 1))].sum()
    (2.0 * rd + (0.5 * ri if with_im else 0.0)).backward()
    assert abs(float(ld.detach()) - float(rd.detach())) < 1e-5 * max(1.0, abs(float(rd.detach())))
    ga, gb = torch.nan_to_num(a_ds.grad), torch.nan_to_num(b_ds.grad)
    assert torch.allclose(ga, gb, rtol=1e-5, atol=1e-7)
    if with_im:
        assert abs(float(li.detach()) - float(ri.detach())) < 1e-5 * float(ri.detach())
        assert torch.allclose(a_im.grad, b_im.grad, rtol=1e-5, atol=1e-7)
