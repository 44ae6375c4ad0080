#!/usr/bin/env python
"""Runs a few single-GPU mapping steps (for an ncu launch list; never a bench value)."""
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
a = ap.parse_args()
dev = torch.device("cuda:0")
r = bench.mapping_bench(dev, 1, False, a.impl, steps=a.steps, warmup=1)
print(r)
