"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b, floor_frac=1e-6):
    """|a-b| / max(|b|, floor_frac * max|b|)  (hard part 3)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    floor = floor_frac * max(np.abs(b).max(), 1e-30)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def l2_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def reference_extension():
    """The UNMODIFIED reference CUDA extension built into baseline/_ref (travels to the GPU box), or None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "diff_gaussian_rasterization")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import diff_gaussian_rasterization as ref
        return ref
    except Exception:
        return None


def _align(off, a=128):
    return (off + a - 1) // a * a


def parse_reference_buffers(geom, binning, img, P, R, W, H):
    """Decode the reference's three scratch byte buffers (GeometryState / BinningState / ImageState
    ::fromChunk, X/cuda_rasterizer/rasterizer_impl.cu:155-194; every chunk 128-B aligned)."""
    import torch
    out = {}
    g = geom.cpu().numpy()
    off = 0

    def take(buf, off, dtype, count):
        off = _align(off)
        n = np.dtype(dtype).itemsize * count
        return np.frombuffer(buf[off:off + n].tobytes(), dtype=dtype), off + n

    out["depths"], off = take(g, off, np.float32, P)
    _, off = take(g, off, np.uint8, 3 * P)           # clamped
    out["internal_radii"], off = take(g, off, np.int32, P)
    m2, off = take(g, off, np.float32, 2 * P)
    out["means2D"] = m2.reshape(P, 2)
    c3, off = take(g, off, np.float32, 6 * P)
    out["cov3D"] = c3.reshape(P, 6)
    co, off = take(g, off, np.float32, 4 * P)
    out["conic_opacity"] = co.reshape(P, 4)
    _, off = take(g, off, np.float32, 3 * P)         # rgb
    out["tiles_touched"], off = take(g, off, np.uint32, P)
    if R > 0:
        b = binning.cpu().numpy()
        off = 0
        out["point_list"], off = take(b, off, np.uint32, R)
        _, off = take(b, off, np.uint32, R)
        out["keys"], off = take(b, off, np.uint64, R)
    i = img.cpu().numpy()
    off = 0
    N = W * H
    ft, off = take(i, off, np.float32, N)
    out["final_T"] = ft.reshape(H, W)
    nc, off = take(i, off, np.uint32, N)
    out["n_contrib"] = nc.reshape(H, W)
    rg, off = take(i, off, np.uint32, 2 * N)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = rg.reshape(N, 2)[:tiles]
    return out
