"""ncu target: the SVGD step at the small BASELINE configs (C2: n=20, D=273,610; C1: n=10, D=501)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import ops
dev = torch.device("cuda", 0)
for n, D in ((20, 273_664), (10, 512), (10, 1_048_576)):
    X = torch.randn(n, D, device=dev) * 0.05
    G = torch.randn(n, D, device=dev) * 1e-3
    out = torch.empty_like(X)
    sc = ops.SvgdScratch.allocate(n, dev)
    for _ in range(3):
        ops.svgd_step(X, G, out, sc, 0.01, 1.0, 768.0)
    torch.cuda.synchronize()
