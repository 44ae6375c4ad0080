"""splatam_b200 -- B200-native (sm_100a) differentiable 3D-Gaussian rasterizer behind SplaTAM's
``GaussianRasterizer`` / ``GaussianRasterizationSettings`` operator API."""
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians,
                         _RasterizeGaussians)
from ._lib import SplatamB200Error, load as load_library

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "SplatamB200Error", "load_library"]
