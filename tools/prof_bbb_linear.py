"""ncu target: a few launches of the fused BBBLinear forward at the CivilComments head shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import bbb_layers, ops  # noqa: E402

dev = torch.device("cuda", 0)
x = torch.randn(16, 768, device=dev)
w_mu = 0.1 * torch.randn(768, 768, device=dev)
w_rho = torch.full((768, 768), -3.0, device=dev)
b_mu, b_rho = torch.zeros(768, device=dev), torch.full((768,), -3.0, device=dev)
for _ in range(4):
    ops.bbb_linear_fwd(x, w_mu, w_rho, b_mu, b_rho, seed=1, stream_id=2, workspace=bbb_layers._workspace)
torch.cuda.synchronize()
