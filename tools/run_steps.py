#!/usr/bin/env python
"""Runs N fwd+bwd raster steps of one workload (for ncu captures; never a bench value)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--workload", default="config3")
ap.add_argument("--gaussians", type=int, default=0)
a = ap.parse_args()
scene, _ = bench.make_scene(a)
dev = torch.device("cuda:0")
Rast, Settings = bench.get_ops(a.impl)
step, _, _, _ = bench.gpu_step_fn(scene, dev, Rast, Settings)
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
print("done", a.impl, a.steps)
