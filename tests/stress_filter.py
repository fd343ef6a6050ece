"""Randomised cross-check of the two neighbourhood-filter kernels (tcgen05 banded-Toeplitz GEMM vs FFMA) on shapes
that take the tensor-core path: K >= 256, D >= 48, at least 32 tiles, any band.  Not collected by pytest.
usage: python tests/stress_filter.py [cases]"""
import math
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rnd = random.Random(7)
lib = somcb._lib.load()
bad = tc_cases = 0
for i in range(cases):
    k = rnd.randint(256, 9000)
    d = rnd.choice([48, 50, 64, 65, 96, 128, 130, 200, 256, 333, 512, 700, 1024])
    rng = math.exp(rnd.uniform(math.log(0.3), math.log(3.0 * k)))
    if lib.som_filter_workspace_bytes(k, d, float(rng)) == 0:
        continue
    tc_cases += 1
    g = torch.Generator(device="cuda").manual_seed(i)
    w = torch.randn(k, d, generator=g, device="cuda") * rnd.choice([1e-3, 1.0, 50.0])
    scale = rnd.choice([1.0, 0.37, 2.0 / (k * d)])
    a = ops.neighbourhood_filter(w, rng, scale=scale)
    b = ops.neighbourhood_filter(w, rng, scale=scale, tensor_cores=False)
    err = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
    mx = float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))
    if not (err <= 3e-6 and mx <= 3e-6) or not bool(torch.isfinite(a).all()):
        bad += 1
        print(f"case {i}: K={k} D={d} range={rng:.3g} scale={scale:.3g}: rel_fro {err:.2e} rel_max {mx:.2e}")
print(f"stress filter: {tc_cases} tensor-core cases of {cases}, {bad} failures")
sys.exit(1 if bad else 0)
