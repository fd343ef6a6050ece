"""Probe one BMU shape on the GPU with a chosen variant: parity vs the CPU oracle + timing.
usage: python tests/tc_probe.py <fmaps> <pH> <K> <variant> [fresh]   (4x32x32 fmaps)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402
from oracle.step_oracle import make_oracle_codebook, synthetic_fmaps, trained_like_codebook  # noqa: E402
from _helpers import assert_bmu_parity, flat_patches  # noqa: E402

b, p, k, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fresh = len(sys.argv) > 5
pd = (p, p)
d = 4 * p * p
x = synthetic_fmaps(b, 4242)
if fresh:
    torch.manual_seed(0)
    w = torch.empty(k, d).uniform_(-1 / k, 1 / k)
else:
    w = trained_like_codebook(k, pd, 7)
scale = float(os.environ.get("SOM_PROBE_SCALE", "1"))      # data and codebook magnitude (exercises the FP16 mode's scaling)
if scale != 1.0:
    x, w = x * scale, w * scale
dev = "cuda:0"
xd, wd = x.to(dev), w.to(dev)
geom = ops.geometry(x.shape, pd)
cn = ops.prepare_codebook(wd)
t0 = time.time()
idx = ops.bmu(xd, geom, wd, cn, variant=variant)
torch.cuda.synchronize()
print(f"shape fmaps={b} P={p} D={d} K={k} variant={variant} first call {time.time() - t0:.3f}s", flush=True)
n_check = min(b, max(1, 65536 // geom[2] // geom[3] * p * p))
oc = make_oracle_codebook(w, pd, (32, 32), 4, k // 2)
with torch.no_grad():
    ref = oc.get_patches_bmu(x[:n_check])
flat = flat_patches(x[:n_check], pd)
got = idx[:flat.shape[0]].cpu()
nbad = -1 if os.environ.get('SOM_PROBE_NOCHECK') else assert_bmu_parity(got, ref, flat, w)
ffma = ops.bmu(xd, geom, wd, cn, variant=ops.SOM_BMU_FFMA)
diff = int((ffma != idx).sum())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    ops.bmu(xd, geom, wd, cn, variant=variant)
e0.record()
reps = int(os.environ.get('SOM_PROBE_REPS', '5'))
for _ in range(reps):
    ops.bmu(xd, geom, wd, cn, variant=variant)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
npat = idx.numel()
try:
    import ctypes
    cyc = (ctypes.c_longlong * 2)()
    lib = somcb._lib.load()
    if lib.som_debug_tc_cycles(cyc) == 0 and cyc[1] > 0 and d <= 16 and variant == 2:
        print(f"  [config S] CTA0 MMA loop: {cyc[0] / cyc[1]:.0f} cycles per 128x256 tile over {cyc[1]} tiles", flush=True)
except Exception as e:  # noqa: BLE001
    print("  (no cycle probe:", e, ")")
print(f"  OK: {nbad} near-tie diffs vs oracle on {flat.shape[0]} patches, {diff} diffs vs FFMA on {npat}; "
      f"{ms:.3f} ms -> {npat / ms * 1e3:.3e} patches/s, {2.0 * k * d * npat / ms / 1e9:.1f} TFLOP/s", flush=True)
