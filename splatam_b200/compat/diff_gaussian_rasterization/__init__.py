"""Drop-in alias: put ``splatam_b200/compat`` on ``sys.path`` (before any install of the reference
extension) and SplaTAM's unmodified ``from diff_gaussian_rasterization import GaussianRasterizer``
(scripts/splatam.py:37, utils/recon_helpers.py:2, utils/eval_helpers.py:17) resolves to the B200 path."""
from splatam_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                     rasterize_gaussians, _RasterizeGaussians)
