"""Test / benchmark infrastructure: imports the UNMODIFIED reference Python (SplaTAM's own `utils/*.py` and
`scripts/splatam.py`) so that tests and `bench.py --impl reference` can run the reference's stock code path.

Nothing under `splatam_b200/` imports this module.  No reference source is copied into the repository: the
files are read where they lie -- `/root/reference` in the build container, or the git-ignored install
`baseline/_ref/SplaTAM/` (written by `__graft_entry__.build_reference()`, which travels to the GPU box next to
the compiled reference extension).

The reference imports packages that are absent from this image (matplotlib, imageio, natsort, kornia,
pytorch_msssim, torchmetrics, its own dataset readers).  None of them is on the path that is exercised here
(`get_loss`, `initialize_optimizer`, `setup_camera`, the slam helpers, `keyframe_selection_overlap`, ATE), so
they are stubbed in `sys.modules` for the duration of the import, as SURVEY.md App. C describes.
"""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref", "SplaTAM"))
_CACHE = {}


def ref_root():
    """Directory holding the reference's `utils/` and `scripts/`, or None."""
    for p in _CANDIDATES:
        if os.path.isfile(os.path.join(p, "utils", "slam_helpers.py")) and \
                os.path.isfile(os.path.join(p, "scripts", "splatam.py")):
            return p
    return None


def available():
    return ref_root() is not None


class _Anything:
    """Stand-in for classes of absent packages that the reference instantiates at import time."""

    def __init__(self, *a, **k):
        pass

    def cuda(self, *a, **k):
        return self

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed dependency of the reference was called")

    def __getattr__(self, name):
        return _Anything()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda attr: _Anything     # any other symbol resolves to the stand-in class
    return m


def _stub_missing():
    names = ["matplotlib", "matplotlib.pyplot", "imageio", "natsort", "kornia", "kornia.geometry",
             "kornia.geometry.linalg", "pytorch_msssim", "torchmetrics", "torchmetrics.image", "torchmetrics.image.lpip",
             "open3d", "wandb", "cv2"]
    added = []
    for n in names:
        if n in sys.modules:
            continue
        try:
            if n in ("wandb", "cv2"):       # present in the image: only stubbed if the real import fails
                importlib.import_module(n)
                continue
        except Exception:
            pass
        if n in ("wandb", "cv2") and n in sys.modules:
            continue
        sys.modules[n] = _stub(n)
        added.append(n)
    return added


def load(rasterizer_pkg):
    """Returns a namespace with the reference's modules bound to `rasterizer_pkg` (a module exporting
    GaussianRasterizer / GaussianRasterizationSettings: the reference extension or splatam_b200.compat's alias):
    .slam_helpers .slam_external .recon_helpers .keyframe_selection .common_utils .splatam (scripts/splatam.py)
    .eval_helpers (ATE / PSNR helpers).  Modules are loaded under private names, once per rasterizer package."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference Python not found (neither /root/reference nor baseline/_ref/SplaTAM)")
    key = (root, rasterizer_pkg.__name__, id(rasterizer_pkg))
    if key in _CACHE:
        return _CACHE[key]
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "utils" or k.startswith("utils.") or k == "datasets" or k.startswith("datasets.")
             or k == "diff_gaussian_rasterization"}
    for k in saved:
        del sys.modules[k]
    added = _stub_missing()
    sys.modules["diff_gaussian_rasterization"] = rasterizer_pkg
    # the dataset readers need imageio / natsort / kornia; only the (pure torch) geometry utilities are wanted
    ds = _stub("datasets"); ds.__path__ = []
    gd = _stub("datasets.gradslam_datasets"); gd.__path__ = [os.path.join(root, "datasets", "gradslam_datasets")]
    sys.modules["datasets"], sys.modules["datasets.gradslam_datasets"] = ds, gd
    sys.path.insert(0, root)
    ns = types.SimpleNamespace(root=root)
    stdout = sys.stdout
    try:
        sys.stdout = open(os.devnull, "w")            # scripts/splatam.py prints sys.path at import
        for name in ("slam_external", "slam_helpers", "recon_helpers", "keyframe_selection", "common_utils",
                     "eval_helpers"):
            setattr(ns, name, importlib.import_module("utils." + name))
        spec = importlib.util.spec_from_file_location("_ref_scripts_splatam", os.path.join(root, "scripts", "splatam.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ns.splatam = mod
    finally:
        sys.stdout.close()
        sys.stdout = stdout
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k == "datasets"
                  or k.startswith("datasets.") or k == "diff_gaussian_rasterization"]:
            del sys.modules[k]
        for k in added:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _CACHE[key] = ns
    return ns


def install(dst=None):
    """Copies the reference's Python tree (scripts/, utils/, datasets/gradslam_datasets/geometryutils.py, configs/)
    UNMODIFIED from /root/reference into baseline/_ref/SplaTAM (git-ignored; travels to the GPU box).  A no-op where
    /root/reference is absent."""
    import shutil
    src = "/root/reference"
    dst = dst or _CANDIDATES[1]
    if not os.path.isdir(os.path.join(src, "utils")):
        return os.path.isdir(os.path.join(dst, "utils"))
    def _writable(func, path, _exc):          # the source tree is read-only and copytree keeps its modes
        os.chmod(path, 0o755)
        func(path)
    for sub in ("utils", "scripts", "configs", os.path.join("datasets", "gradslam_datasets")):
        d = os.path.join(dst, sub)
        if os.path.isdir(d):
            for root, dirs, files in os.walk(d):
                os.chmod(root, 0o755)
            shutil.rmtree(d, onerror=_writable)
        shutil.copytree(os.path.join(src, sub), d)
        for root, dirs, files in os.walk(d):          # keep the install removable / re-installable
            os.chmod(root, 0o755)
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)
    return True
