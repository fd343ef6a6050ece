"""Randomised cross-check of the tensor-core BMU kernels against the FFMA kernel on many small geometries
(not collected by pytest; run by hand on a GPU box: python tests/stress_bmu.py [n_cases] [seed])."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = "cuda:0"
bad_cases = 0
for case in range(n_cases):
    c = rnd.choice([1, 2, 3, 4, 5])
    ph, pw = rnd.choice([1, 2, 3, 4, 8, 16]), rnd.choice([1, 2, 3, 4, 8, 16])
    gh, gw = rnd.randint(1, 4), rnd.randint(1, 4)
    h, w = ph * gh, pw * gw
    n = rnd.choice([1, 2, 3, 7, 33, 130, 300] if os.environ.get("SOM_STRESS_BIG") != "1" else [500, 3000, 9000])
    k = rnd.choice([1, 2, 7, 8, 9, 255, 256, 257, 1000, 3000])
    d = c * ph * pw
    g = torch.Generator(device=dev).manual_seed(case)
    x = torch.tanh(torch.randn(n, c, h, w, generator=g, device=dev))
    wgt = torch.tanh(torch.randn(k, d, generator=g, device=dev))
    if k > 4:                                   # plant exact duplicates: the lower index must win
        wgt[k - 1] = wgt[0]
        wgt[k // 2] = wgt[1]
    geom = ops.geometry(x.shape, (ph, pw))
    cn = ops.prepare_codebook(wgt)
    a = ops.bmu(x, geom, wgt, cn, variant=ops.SOM_BMU_TC3X)
    b = ops.bmu(x, geom, wgt, cn, variant=ops.SOM_BMU_FFMA)
    diff = torch.nonzero(a != b).flatten()
    ok = True
    if diff.numel():
        flat = somcb.patchify(x, (ph, pw)).reshape(-1, d).double()
        da = (flat[diff] - wgt[a[diff]].double()).norm(dim=1)
        db = (flat[diff] - wgt[b[diff]].double()).norm(dim=1)
        # equal within the north-star epsilon, or within the fp32 resolution of the expanded form
        # ||c||^2 - 2 x.c itself (a few ulp of ||x||^2 + ||c||^2: dense units at tiny D, where the reference's
        # own cdist cannot separate the two candidates either)
        scale = (flat[diff] ** 2).sum(dim=1) + (wgt[b[diff]].double() ** 2).sum(dim=1)
        ok = bool((((da - db).abs() <= 1e-6 * db.clamp_min(1e-30)) |
                   ((da * da - db * db).abs() <= 5e-7 * scale)).all())
        # duplicates: equal rows must resolve to the lower index in BOTH variants
        same = (wgt[a[diff]] == wgt[b[diff]]).all(dim=1)
        ok = ok and not bool(same.any())
    if int(a.min()) < 0 or int(a.max()) >= k:
        ok = False
    if not ok:
        bad_cases += 1
        print(f"FAIL case {case}: n={n} C={c} H={h} W={w} P={ph}x{pw} D={d} K={k}: {diff.numel()} diffs", flush=True)
torch.cuda.synchronize()
print(f"stress: {n_cases} cases, {bad_cases} failures")
sys.exit(1 if bad_cases else 0)
