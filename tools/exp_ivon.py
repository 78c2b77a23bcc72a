"""Per-call device time of ivon_update over successive steps (is the time data dependent?)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
D = 66_955_072
mean = torch.randn(D, device=dev, generator=g) * 0.05
prec = torch.rand(D, device=dev, generator=g) * 1e-4 + 10.0 / 269038
mom = torch.randn(D, device=dev, generator=g) * 1e-4
dsum = torch.randn(D, device=dev, generator=g) * 0.3
acc = torch.randn(D, device=dev, generator=g) * 2e-5
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
ev[0].record()
for r in range(40):
    ops.ivon_update(acc, dsum, mean, mom, prec, mc_samples=2, step=101 + r, lr=1e-5, beta1=0.9, beta2=0.999,
                    prior_prec=10.0, n_eff=269038.0, tempering=1.0, damping=1e-3)
    ev[r + 1].record()
torch.cuda.synchronize()
print("per-call ms:", " ".join(f"{ev[i].elapsed_time(ev[i+1]):.3f}" for i in range(40)))
print("prec min/max", prec.min().item(), prec.max().item(), "nan", torch.isnan(prec).sum().item(),
      "mom denormal frac", ((mom.abs() < 1.18e-38) & (mom != 0)).float().mean().item(),
      "mean absmax", mean.abs().max().item())
