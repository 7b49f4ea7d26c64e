"""Timing of the empirical-Bayes leg alone (bench.py: eb_leg) — for A/B runs of the likelihood-matrix / dhz kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gync_b200 import lib as L
lib = L.load()
L.check(lib.gync_init(0))
dev = torch.device("cuda", 0)
print(json.dumps(bench.eb_leg(lib, L, torch, dev, torch.cuda.current_stream().cuda_stream), indent=1))
