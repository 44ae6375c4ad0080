"""GPU parity tests proper: the sm_100a path (through the C-ABI) against (1) the CPU oracle,
(2) the committed golden vectors, (3) the unmodified reference extension on the same device, and
(4) size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): tile-ID/depth sort keys bit-identical; colour / depth /
silhouette / gradients within 1e-4 relative float32.
"""
import glob
import os

import numpy as np
import pytest
import torch

import scenes
from util import GOLDEN, l2_rel, reference_extension, rel_err

pytestmark = pytest.mark.gpu

GRADS = ["means3D", "means2D", "colors", "opacities", "scales", "rotations"]
REL = 1e-4


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _check_against(ours, ref, exact_images, tag, grad_tol=REL):
    vis = ref["radii"] > 0
    assert np.array_equal(ours["radii"], ref["radii"]), tag
    assert np.array_equal(ours["tiles_touched"], ref["tiles_touched"]), tag
    assert np.array_equal(_bits(ours["depths"][vis]), _bits(ref["depths"][vis])), tag
    assert np.array_equal(_bits(ours["means2D"][vis]), _bits(ref["means2D"][vis])), tag
    assert np.array_equal(_bits(ours["conic_opacity"][vis]), _bits(ref["conic_opacity"][vis])), tag
    assert ours["num_rendered"] == len(ref["keys"]), tag
    assert np.array_equal(ours["keys"], ref["keys"]), tag + ": sorted (tile|depth) keys must be bit-identical"
    assert np.array_equal(ours["point_list"], ref["point_list"]), tag
    assert np.array_equal(ours["ranges"], ref["ranges"]), tag
    if exact_images:
        assert np.array_equal(ours["n_contrib"], ref["n_contrib"]), tag
        assert np.array_equal(_bits(ours["final_T"]), _bits(ref["final_T"])), tag
        assert np.array_equal(_bits(ours["color"]), _bits(ref["color"])), tag
        assert np.array_equal(_bits(ours["depth"]), _bits(ref["depth"])), tag
    else:  # CPU expf differs from libdevice expf by an ulp or two
        assert (ours["n_contrib"] != ref["n_contrib"]).mean() < 1e-3, tag
        assert (rel_err(ours["color"], ref["color"], 1e-3) > REL).mean() < 1e-3, tag
        assert (rel_err(ours["final_T"], ref["final_T"], 1e-3) > REL).mean() < 1e-3, tag
    for k in GRADS:
        a, b = ours["grad_" + k].reshape(-1), ref["grad_" + k].reshape(-1)
        if k == "rotations" and np.abs(b).max() < 1e-6 * np.abs(ref["grad_scales"]).max():
            # isotropic scene: the quaternion gradient is mathematically zero and pure cancellation noise (~1e-10) in
            # every implementation (the reference's two runs differ by ~100 % here); only its smallness is meaningful
            assert np.abs(a).max() < 1e-6 * np.abs(ref["grad_scales"]).max(), (tag, k)
            continue
        assert l2_rel(a, b) < grad_tol, (tag, k, l2_rel(a, b))
        # element-wise: entries that are sums with heavy cancellation carry float32 summation-order noise
        # (the reference's own atomics are order-nondeterministic), so floor at 1e-3 of the largest entry
        frac = (rel_err(a, b, 1e-3) > 10 * grad_tol).mean()
        assert frac <= max(2e-3, 2.0 / a.size), (tag, k, frac)


SMALL = [scenes.config1, scenes.edge_cases, lambda: scenes.dense_opaque(P=600, w=64, h=48),
         lambda: scenes.config1(seed=11, P=1000, w=200, h=120, bg=(0.3, 0.1, 0.9))]


@pytest.mark.parametrize("make", SMALL)
def test_cuda_vs_oracle(make, cuda_device):
    from gpu_harness import random_dL, run_ours
    sc = make()
    dL = random_dL(sc)
    ours = run_ours(sc, dL)
    o = sc.oracle()
    geo, b, r = o.geometry(), o.binning(), o.render()
    g = o.backward(dL)
    ref = dict(radii=geo["radii"], tiles_touched=geo["tiles_touched"], depths=geo["depths"], means2D=geo["means2D"],
               conic_opacity=geo["conic_opacity"], keys=b["keys"], point_list=b["point_list"], ranges=b["ranges"],
               n_contrib=r["n_contrib"], final_T=r["final_T"], color=r["color"], depth=r["depth"])
    ref.update({"grad_" + k: g[k] for k in GRADS})
    _check_against(ours, ref, exact_images=False, tag=sc.name + " vs oracle")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))) or [None])
def test_cuda_vs_golden(path, cuda_device):
    if path is None:
        pytest.skip("no golden fixtures committed yet")
    from gpu_harness import run_ours
    z = np.load(path)
    sc = scenes.Scene(str(z["name"]), int(z["w"]), int(z["h"]), float(z["fx"]), float(z["fy"]), float(z["cx"]),
                      float(z["cy"]), torch.from_numpy(z["means3D"]), torch.from_numpy(z["colors"]),
                      torch.from_numpy(z["opacities"]), torch.from_numpy(z["scales"]), torch.from_numpy(z["rotations"]),
                      w2c=torch.from_numpy(z["w2c"]), bg=tuple(z["bg"].tolist()))
    ours = run_ours(sc, z["dL_dcolor"])
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    _check_against(ours, ref, exact_images=True, tag=os.path.basename(path))


BIG = [("config1", scenes.config1), ("edge", scenes.edge_cases), ("dense", lambda: scenes.dense_opaque()),
       ("room50k", lambda: scenes.room(P=50_000)), ("tum_aniso", lambda: scenes.room(seed=9, P=200_000, cam=scenes.TUM_FR1, anisotropic=True)),
       ("config3_1M", lambda: scenes.config3()),
       ("splatam816k", lambda: scenes.view_filling(seed=12)),            # one Gaussian per pixel: every Gaussian in view
       ("replica50k", lambda: scenes.view_filling(seed=12, P=50_000, cover=True)),
       ("tum3m_aniso", lambda: scenes.view_filling(seed=15, P=3_000_000, cam=scenes.TUM_FR1, anisotropic=True))]


@pytest.mark.parametrize("name,make", BIG)
def test_cuda_vs_reference_extension(name, make, cuda_device):
    """Same inputs through the unmodified reference extension on the same GPU."""
    if reference_extension() is None:
        pytest.skip("baseline/_ref not built")
    from gpu_harness import random_dL, run_ours, run_ref
    sc = make()
    dL = random_dL(sc)
    ours, ref = run_ours(sc, dL), run_ref(sc, dL)
    ref2 = run_ref(sc, dL)  # the reference's own run-to-run atomics noise calibrates the gradient bound
    noise = max(l2_rel(ref2["grad_" + k], ref["grad_" + k]) for k in GRADS)
    _check_against(ours, ref, exact_images=True, tag=name, grad_tol=max(REL, 4 * noise))


def _err_stats(a, b):
    """Element-wise relative error with the SURVEY floor (1e-6 of the largest entry) + L2-relative error."""
    e = rel_err(a, b, 1e-6).reshape(-1)
    return dict(l2_rel=l2_rel(a, b), p50=float(np.percentile(e, 50)), p99=float(np.percentile(e, 99)),
                p999=float(np.percentile(e, 99.9)), max=float(e.max()), frac_gt_1e4=float((e > 1e-4).mean()),
                frac_gt_1e3=float((e > 1e-3).mean()))


def test_gradients_vs_double_oracle_full_size(cuda_device):
    """BASELINE config[2] (1M Gaussians, 1200x680): all six gradients against the C oracle, whose backward
    accumulates in DOUBLE and is deterministic, element by element.  The same statistics are taken for the reference
    extension (float atomics) so the numbers can be read side by side; everything measured is written to
    gpurun_out/r02_gradient_parity.json (committed under profiles/)."""
    import json
    from gpu_harness import random_dL, run_ours, run_ref
    sc = scenes.config3()
    dL = random_dL(sc)
    ours = run_ours(sc, dL, intermediates=False)
    o = sc.oracle()
    o.render()
    og = o.backward(dL)
    report = {"workload": "config3 1M Gaussians 1200x680", "floor": "1e-6 x max|reference entry|", "ours_vs_oracle": {},
              "reference_vs_oracle": {}, "reference_run_to_run": {}}
    have_ref = reference_extension() is not None
    if have_ref:
        r1, r2 = run_ref(sc, dL), run_ref(sc, dL)
    for k in GRADS:
        if k == "rotations":      # isotropic scene: mathematically zero, float noise in every implementation
            continue
        report["ours_vs_oracle"][k] = _err_stats(ours["grad_" + k], og[k])
        if have_ref:
            report["reference_vs_oracle"][k] = _err_stats(r1["grad_" + k], og[k])
            report["reference_run_to_run"][k] = _err_stats(r2["grad_" + k], r1["grad_" + k])
    os.makedirs(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "r02_gradient_parity.json"), "w") as f:
        json.dump(report, f, indent=1)
    for k, st in report["ours_vs_oracle"].items():
        assert st["l2_rel"] < REL, (k, st)
        # element-wise: float32 sums with cancellation cannot hold 1e-4 on every entry (the reference's own atomics
        # do not either); the bulk must, and the tail must not be worse than the reference's tail
        assert st["p50"] < REL and st["frac_gt_1e3"] < 2e-3, (k, st)
        if have_ref:
            rs = report["reference_vs_oracle"][k]
            assert st["p99"] <= max(2.0 * rs["p99"], REL), (k, st, rs)


def test_full_size_properties(cuda_device):
    """Size-independent properties at BASELINE config[2] (1M Gaussians, 1200x680)."""
    from gpu_harness import random_dL, run_ours
    sc = scenes.config3()
    dL1, dL2 = random_dL(sc, 3), random_dL(sc, 4)
    a = run_ours(sc, dL1)
    keys, lst = a["keys"], a["point_list"]
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert np.all(lst[1:][same] > lst[:-1][same])
    assert int(a["tiles_touched"].sum()) == a["num_rendered"] == int((a["ranges"][:, 1] - a["ranges"][:, 0]).sum())
    assert np.array_equal((keys & 0xFFFFFFFF).astype(np.uint32), _bits(a["depths"])[lst])
    # forward is deterministic; backward is linear in dL/dcolor
    b = run_ours(sc, dL2, intermediates=False)
    c = run_ours(sc, dL1 + dL2, intermediates=False)
    assert np.array_equal(_bits(a["color"]), _bits(b["color"]))
    for k in GRADS:
        if k == "rotations":   # isotropic scene: the quaternion gradient is pure cancellation noise (~1e-12)
            assert np.abs(c["grad_" + k]).max() < 1e-6 * np.abs(c["grad_scales"]).max()
            continue
        assert l2_rel(c["grad_" + k], a["grad_" + k] + b["grad_" + k]) < 1e-5, k
    # a zero upstream gradient gives exactly zero gradients; transmittance is in (0, 1]
    z = run_ours(sc, np.zeros_like(dL1), intermediates=False)
    assert all(not z["grad_" + k].any() for k in GRADS)
    assert a["final_T"].min() >= 0 and a["final_T"].max() <= 1.0


def test_api_edge_cases(cuda_device):
    import splatam_b200 as S
    dev = cuda_device
    sc = scenes.config1()
    rs = sc.settings(S.GaussianRasterizationSettings, dev)
    rast = S.GaussianRasterizer(rs)
    inp = sc.inputs(dev)
    with pytest.raises(Exception, match="excatly one of either SHs"):
        rast(means3D=inp["means3D"], means2D=inp["means2D"], opacities=inp["opacities"], scales=inp["scales"],
             rotations=inp["rotations"])
    with pytest.raises(Exception, match="scale/rotation pair"):
        rast(means3D=inp["means3D"], means2D=inp["means2D"], opacities=inp["opacities"],
             colors_precomp=inp["colors_precomp"])
    # empty scene: background-free zero images like the reference (rasterize_points.cu:67-81)
    e = {k: v[:0] for k, v in inp.items()}
    color, radii, depth = rast(**e)
    assert color.shape == (3, sc.h, sc.w) and radii.numel() == 0 and not color.any()
    # everything behind the camera: R == 0, background only, depth default 15
    inp2 = sc.inputs(dev)
    inp2["means3D"][:, 2] = -1.0
    rs2 = rs._replace(bg=torch.tensor([0.25, 0.5, 0.75], device=dev))
    color, radii, depth = S.GaussianRasterizer(rs2)(**inp2)
    assert not radii.any() and torch.allclose(color[1], torch.tensor(0.5, device=dev)) and torch.all(depth == 15.0)
    # non-contiguous inputs (SplaTAM passes a strided view, utils/slam_helpers.py:288)
    inp3 = sc.inputs(dev)
    m4 = torch.cat([inp3["means3D"], torch.ones(sc.P, 1, device=dev)], 1)
    inp3["means3D"] = (torch.eye(4, device=dev) @ m4.T).T[:, :3]
    c3, _, _ = rast(**inp3)
    c0, _, _ = rast(**sc.inputs(dev))
    assert torch.equal(c3, c0)
    # markVisible
    vis = rast.markVisible(inp2["means3D"])
    assert vis.dtype == torch.bool and not vis.any()
    assert rast.markVisible(sc.inputs(dev)["means3D"]).all()


def test_cov3d_precomp_branch_vs_reference(cuda_device):
    """The cov3D_precomp branch of the operator API (X/cuda_rasterizer/forward.cu:203-214,
    backward.cu:393-395): same images as the scale/rotation path, dL/dcov3D against the reference."""
    ref = reference_extension()
    import splatam_b200 as S
    dev = cuda_device
    sc = scenes.config1(seed=13, P=2000, w=160, h=96)
    cov = torch.from_numpy(sc.oracle().geometry()["cov3D"]).to(dev)   # bit-exact Sigma of the scale/rot path
    dL = torch.randn(3, sc.h, sc.w, generator=torch.Generator().manual_seed(5)).to(dev)

    def run(mod):
        rs = sc.settings(mod.GaussianRasterizationSettings, dev)
        inp = sc.inputs(dev, requires_grad=True)
        c = cov.clone().requires_grad_(True)
        color, radii, depth = mod.GaussianRasterizer(rs)(means3D=inp["means3D"], means2D=inp["means2D"],
                                                         opacities=inp["opacities"], colors_precomp=inp["colors_precomp"],
                                                         cov3D_precomp=c)
        color.backward(dL)
        return color.detach().cpu().numpy(), radii.cpu().numpy(), c.grad.cpu().numpy(), inp["means3D"].grad.cpu().numpy()

    col, rad, gcov, gmean = run(S)
    # identical to the scale/rotation path
    c0, r0, _ = S.GaussianRasterizer(sc.settings(S.GaussianRasterizationSettings, dev))(**sc.inputs(dev))
    assert np.array_equal(col.view(np.uint32), c0.cpu().numpy().view(np.uint32)) and np.array_equal(rad, r0.cpu().numpy())
    if ref is not None:
        rcol, rrad, rgcov, rgmean = run(ref)
        assert np.array_equal(col.view(np.uint32), rcol.view(np.uint32)) and np.array_equal(rad, rrad)
        assert l2_rel(gcov, rgcov) < REL and l2_rel(gmean, rgmean) < REL


@pytest.mark.parametrize("make", [scenes.config1, scenes.edge_cases, lambda: scenes.dense_opaque(),
                                  lambda: scenes.room(seed=9, P=100_000, cam=scenes.TUM_FR1, anisotropic=True)])
def test_fused_two_set_render_equals_two_passes(make, cuda_device):
    """N1: the fused RGB + depth/silhouette render against two plain renders over the same geometry."""
    import splatam_b200 as S
    dev = cuda_device
    sc = make()
    rs = sc.settings(S.GaussianRasterizationSettings, dev)
    rast = S.GaussianRasterizer(rs)
    g = torch.Generator().manual_seed(8)
    extra = torch.rand(sc.P, 3, generator=g).to(dev) * 3.0
    dL1 = torch.randn(3, sc.h, sc.w, generator=g).to(dev)
    dL2 = torch.randn(3, sc.h, sc.w, generator=g).to(dev)
    # two plain passes
    a = sc.inputs(dev, requires_grad=True)
    ex_a = extra.clone().requires_grad_(True)
    m2b = torch.zeros_like(a["means3D"], requires_grad=True)
    c1, r1, d1 = rast(**a)
    c2, _, _ = rast(means3D=a["means3D"], means2D=m2b, opacities=a["opacities"], colors_precomp=ex_a,
                    scales=a["scales"], rotations=a["rotations"])
    ((c1 * dL1).sum() + (c2 * dL2).sum()).backward()
    # fused
    b = sc.inputs(dev, requires_grad=True)
    ex_b = extra.clone().requires_grad_(True)
    f1, f2, rf, df = rast.forward_fused(means3D=b["means3D"], means2D=b["means2D"], opacities=b["opacities"],
                                        colors_precomp=b["colors_precomp"], colors_extra=ex_b, scales=b["scales"],
                                        rotations=b["rotations"])
    ((f1 * dL1).sum() + (f2 * dL2).sum()).backward()
    assert torch.equal(f1, c1) and torch.equal(f2, c2) and torch.equal(rf, r1) and torch.equal(df, d1)
    for k in ["means3D", "colors_precomp", "opacities", "scales", "rotations"]:
        ga, gb = a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy()
        if k == "rotations" and np.abs(ga).max() < 1e-9:
            continue
        assert l2_rel(gb, ga) < 2e-5, (k, l2_rel(gb, ga))
    assert l2_rel(ex_b.grad.cpu().numpy(), ex_a.grad.cpu().numpy()) < 2e-5
    # means2D sink: first colour set only
    assert l2_rel(b["means2D"].grad.cpu().numpy(), a["means2D"].grad.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("deg", [0, 1, 3])
def test_sh_colour_branch_vs_reference(deg, cuda_device):
    """The shs / sh_degree branch of the operator (unused by SplaTAM) against the reference extension."""
    ref = reference_extension()
    if ref is None:
        pytest.skip("baseline/_ref not built")
    import splatam_b200 as S
    dev = cuda_device
    sc = scenes.config1(seed=17, P=3000, w=160, h=96)
    g = torch.Generator().manual_seed(deg)
    M = 16
    shs = (torch.randn(sc.P, M, 3, generator=g) * 0.4).to(dev)
    dL = torch.randn(3, sc.h, sc.w, generator=g).to(dev)

    def run(mod):
        rs = sc.settings(mod.GaussianRasterizationSettings, dev)._replace(sh_degree=deg, campos=torch.tensor([0.1, -0.2, 0.05], device=dev))
        inp = sc.inputs(dev, requires_grad=True)
        sh = shs.clone().requires_grad_(True)
        color, radii, depth = mod.GaussianRasterizer(rs)(means3D=inp["means3D"], means2D=inp["means2D"], shs=sh,
                                                         opacities=inp["opacities"], scales=inp["scales"],
                                                         rotations=inp["rotations"])
        color.backward(dL)
        return (color.detach().cpu().numpy(), radii.cpu().numpy(), sh.grad.cpu().numpy(), inp["means3D"].grad.cpu().numpy(),
                inp["opacities"].grad.cpu().numpy())

    ours, theirs = run(S), run(ref)
    assert np.array_equal(ours[1], theirs[1])
    assert (rel_err(ours[0], theirs[0], 1e-3) > REL).mean() < 1e-4
    for a, b, name in zip(ours[2:], theirs[2:], ["shs", "means3D", "opacities"]):
        assert l2_rel(a, b) < REL, (name, l2_rel(a, b))
    assert not ours[2][:, (deg + 1) ** 2:, :].any(), "coefficients above sh_degree get no gradient"


def test_sync_free_forward_matches_and_graph_captures(cuda_device):
    """max_rendered (sync-free, fixed-capacity) mode: identical images, gradients within atomics noise, overflow
    flag when the capacity is too small, and forward+backward replayable from a CUDA graph."""
    import splatam_b200 as S
    dev = cuda_device
    sc = scenes.config3(P=200_000)
    rs = sc.settings(S.GaussianRasterizationSettings, dev)
    dL = torch.randn(3, sc.h, sc.w, generator=torch.Generator().manual_seed(3)).to(dev)
    a = sc.inputs(dev, requires_grad=True)
    c0, r0, d0 = S.GaussianRasterizer(rs)(**a)
    R = c0.grad_fn.state.num_rendered
    c0.backward(dL)
    rast = S.GaussianRasterizer(rs, max_rendered=int(R * 1.3))
    b = sc.inputs(dev, requires_grad=True)
    c1, r1, d1 = rast(**b)
    c1.backward(dL)
    assert rast.last_counts() == (R, False)
    assert torch.equal(c1, c0) and torch.equal(r1, r0) and torch.equal(d1, d0)
    for k in ["means3D", "colors_precomp", "opacities", "scales"]:
        assert l2_rel(b[k].grad.cpu().numpy(), a[k].grad.cpu().numpy()) < 1e-5, k
    # overflow is reported, not silent
    small = S.GaussianRasterizer(rs, max_rendered=R // 2)
    with torch.no_grad():
        small(**sc.inputs(dev))
    assert small.last_counts() == (R, True)
    # CUDA graph: capture forward + backward once, replay with new inputs in the static buffers
    static = sc.inputs(dev, requires_grad=True)
    leaves = [static[k] for k in ["means3D", "means2D", "colors_precomp", "opacities", "scales", "rotations"]]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for t in leaves:
                t.grad = None
            rast(**static)[0].backward(dL)
    torch.cuda.current_stream().wait_stream(side)
    for t in leaves:
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, _, _ = rast(**static)
        out.backward(dL)
    with torch.no_grad():
        static["colors_precomp"].mul_(0.5)           # change an input in place, replay
    graph.replay()
    torch.cuda.synchronize()
    ref_in = sc.inputs(dev, requires_grad=True)
    with torch.no_grad():
        ref_in["colors_precomp"].mul_(0.5)
    cr, _, _ = S.GaussianRasterizer(rs)(**ref_in)
    cr.backward(dL)
    assert torch.equal(out, cr)
    assert l2_rel(static["means3D"].grad.cpu().numpy(), ref_in["means3D"].grad.cpu().numpy()) < 1e-5


def test_more_than_65535_tiles_uses_32bit_tile_keys(cuda_device):
    """4112x4112 -> 257x257 = 66049 tiles: the tile-id sort switches to 32-bit keys (17 bits, as the reference's
    getHigherMsb gives); checked against the oracle in the synchronous and the sync-free mode."""
    import splatam_b200 as S
    from gpu_harness import random_dL, run_ours
    sc = scenes.config1(seed=21, P=1500, w=4112, h=4112)
    ours = run_ours(sc, random_dL(sc, 1))
    o = sc.oracle()
    geo, b, r = o.geometry(), o.binning(), o.render()
    assert np.array_equal(ours["radii"], geo["radii"])
    assert np.array_equal(ours["keys"], b["keys"]) and np.array_equal(ours["point_list"], b["point_list"])
    assert np.array_equal(ours["ranges"], b["ranges"])
    assert (ours["n_contrib"] != r["n_contrib"]).mean() < 1e-4
    assert (rel_err(ours["color"], r["color"], 1e-3) > REL).mean() < 1e-4
    dev = cuda_device
    rs = sc.settings(S.GaussianRasterizationSettings, dev)
    with torch.no_grad():
        c1, _, _ = S.GaussianRasterizer(rs, max_rendered=ours["num_rendered"] + 1000)(**sc.inputs(dev))
    assert S.GaussianRasterizer.last_counts() == (ours["num_rendered"], False)
    assert np.array_equal(c1.cpu().numpy().view(np.uint32), ours["color"].view(np.uint32))
