"""Direct-LDG vs TMA-staged elementwise kernels at the SURVEY §8d sizes (tuning aid, not a benchmark)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beyond_deep_ensembles_b200 import _lib, ops

dev = torch.device("cuda", 0)
h = _lib.get()
g = torch.Generator(device=dev).manual_seed(1)


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, nbytes, fn, reset=None):
    for v in (1, 2):
        h.bde_tune(b"ew_variant", v)
        if reset:
            reset()
        ms = timeit(fn)
        print(f"{name:24s} variant={v} {ms:.4f} ms {nbytes / ms / 1e6:8.0f} GB/s", flush=True)
    h.bde_tune(b"ew_variant", 0)


D, K = 23_880_960, 10
theta = torch.randn(D, device=dev, generator=g) * 0.05
mean = theta + torch.randn(D, device=dev, generator=g) * 0.01
sq = mean * mean + 1e-4
ring = torch.randn(K, D, device=dev, generator=g) * 0.01
u = [0]
def swag_upd():
    u[0] += 1
    ops.swag_update(theta, mean, sq, ring[(u[0] - 1) % K], u[0])
report("swag_update 23.9M", 24 * D, swag_upd)
del theta, mean, sq, ring
D = 66_955_072
mean0 = torch.randn(D, device=dev, generator=g) * 0.05
prec0 = torch.rand(D, device=dev, generator=g) * 1e-4 + 10.0 / 269038
mom0 = torch.randn(D, device=dev, generator=g) * 1e-4
dsum0 = torch.randn(D, device=dev, generator=g) * 0.3
acc0 = torch.randn(D, device=dev, generator=g) * 2e-5
mean, prec, mom, dsum, acc = (t.clone() for t in (mean0, prec0, mom0, dsum0, acc0))
theta = torch.zeros(D, device=dev)
grad = torch.randn(D, device=dev, generator=g) * 1e-5
def reset():
    for a, b in ((mean, mean0), (prec, prec0), (mom, mom0), (dsum, dsum0), (acc, acc0)):
        a.copy_(b)
report("ivon_sample philox", 20 * D, lambda: ops.ivon_sample(mean, prec, dsum, theta, first=False, seed=1, stream_id=3, n_eff=269038.0), reset)
report("ivon_sample first", 16 * D, lambda: ops.ivon_sample(mean, prec, dsum, theta, first=True, seed=1, stream_id=3, n_eff=269038.0), reset)
report("ivon_sample injected", 24 * D, lambda: ops.ivon_sample(mean, prec, dsum, theta, first=False, eps=grad, n_eff=269038.0), reset)
report("ivon_accumulate", 12 * D, lambda: ops.ivon_accumulate(acc, grad, first=False), reset)
st = [100]
def upd():
    st[0] += 1
    ops.ivon_update(acc, dsum, mean, mom, prec, mc_samples=2, step=st[0], lr=1e-5, beta1=0.9, beta2=0.999,
                    prior_prec=10.0, n_eff=269038.0, tempering=1.0, damping=1e-3)
report("ivon_update", 32 * D, upd, reset)
a, b = mean, theta
print("torch copy", 8 * D / timeit(lambda: b.copy_(a)) / 1e6, "GB/s")
