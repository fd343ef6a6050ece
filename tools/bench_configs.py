"""Per-config device timings of the SOM hot path on ONE B200: every BASELINE.json config at its
per-GPU shape -- BMU alone, and (training configs) each kernel of the fused step + the whole step.

    python tools/bench_configs.py [--out gpurun_out/configs.json] [--only C3,C5]

CUDA events on the current stream, >= 3 warm-ups, inputs resident in HBM.  Roofline fractions use
MEASURED_PEAKS.json (HBM copy GB/s; sustained bf16 / 2 / 3 for the fp32-faithful 3xTF32 tensor roof).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402
from somcb import ops  # noqa: E402

CONFIGS = {
    # name: fmaps per step per GPU, patch, units per GPU, neighbourhood range, train?
    "C1": dict(fmaps=8, p=(4, 4), K=1024, rng=512, train=True,
               note="train_codebook.py shape: 512 patches/step, D=64 (launch-latency bound: us/step)"),
    "C2": dict(fmaps=39063, p=(2, 2), K=4096, rng=2048, train=False,
               note="fine-patch tokenisation: 10 000 128 patches, D=16, BMU only"),
    "C3": dict(fmaps=4096, p=(32, 32), K=512, rng=256, train=True,
               note="whole-fmap codebook: 4096 patches, D=4096"),
    "C4": dict(fmaps=16384, p=(4, 4), K=16384, rng=8192, train=True,
               note="data-parallel SOM: 1 048 576 patches/GPU/step, D=64"),
    "C5": dict(fmaps=65536, p=(8, 8), K=32768, rng=16384, train=False,
               note="unit-sharded search, one GPU's shard: 1 048 576 patches, D=256, 32 768 of 262 144 units"),
}


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
    hbm = peaks["hbm_gbs"]
    tc_roof = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) / 6.0
    dev = torch.device("cuda", 0)
    lib = somcb._lib.load()
    names = [n for n in CONFIGS if not args.only or n in args.only.split(",")]
    rows = {}
    for name in names:
        c = CONFIGS[name]
        k, pd = c["K"], c["p"]
        d = 4 * pd[0] * pd[1]
        g = torch.Generator(device=dev).manual_seed(123)
        x = torch.empty(c["fmaps"], 4, 32, 32, device=dev)
        for lo in range(0, c["fmaps"], 8192):
            hi = min(c["fmaps"], lo + 8192)
            x[lo:hi] = torch.tanh(torch.randn(hi - lo, 4, 32, 32, generator=g, device=dev))
        geom = ops.geometry(x.shape, pd)
        n = ops.n_patches_of(geom)
        # trained-like codebook: K distinct data patches of an independent pool
        gp = torch.Generator(device=dev).manual_seed(7)
        pool = torch.tanh(torch.randn(max(8, (k * d) // 4096 + 1), 4, 32, 32, generator=gp, device=dev))
        w = somcb.patchify(pool, pd).reshape(-1, d)[:k].contiguous()
        assert w.shape == (k, d)
        cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                            init_neighbour_range=c["rng"]).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(w)
        wd = cb.codebook.weight.data
        cn = ops.prepare_codebook(wd)
        variant = lib.som_bmu_pick_variant(n, d, k)
        reps = args.reps if n * k * d < 5e12 else 3
        row = {"note": c["note"], "patches": n, "D": d, "K": k, "variant": int(variant)}
        ms = timed(lambda: ops.bmu(x, geom, wd, cn), reps)
        fl = 2.0 * k * d * n
        row["bmu"] = {"ms": ms, "patches_per_s": n / ms * 1e3, "tflops": fl / ms / 1e9,
                      "frac_3xtf32_roof": fl / ms / 1e9 / tc_roof,
                      "hbm_frac": (4 * d + 8) * n / ms / 1e6 / hbm}
        if c["train"]:
            rng = c["rng"]
            tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9)
            bmu = ops.bmu(x, geom, wd, cn)
            wt = ops.neighbourhood_filter(wd, rng)
            rbar, _, _ = ops.accumulate(x, geom, bmu, wt, k, want_sse=True)
            grad = ops.neighbourhood_filter(rbar, rng, scale=2.0 / x.numel())
            m, v = torch.zeros_like(wd), torch.zeros_like(wd)
            w2 = wd.clone()
            band = min(k, 2 * int(6.72 * rng ** 0.5) + 1)
            t_f = timed(lambda: ops.neighbourhood_filter(wd, rng), reps)
            t_a = timed(lambda: ops.accumulate(x, geom, bmu, wt, k, want_sse=True), reps)
            t_ad = timed(lambda: ops.adam_step(w2, m, v, grad, 1e-4, 1), reps)
            t_n = timed(lambda: ops.prepare_codebook(wd), reps)
            row["filter"] = {"ms": t_f, "band": band, "tflops_ffma": 2.0 * k * d * band / t_f / 1e9,
                             "hbm_frac": 8.0 * k * d / t_f / 1e6 / hbm}
            row["accumulate"] = {"ms": t_a, "hbm_gbs": ((4 * d + 8) * n + 8.0 * k * d) / t_a / 1e6,
                                 "hbm_frac": ((4 * d + 8) * n + 8.0 * k * d) / t_a / 1e6 / hbm}
            row["adam"] = {"ms": t_ad, "hbm_frac": 28.0 * k * d / t_ad / 1e6 / hbm}
            row["norm2"] = {"ms": t_n}
            t_s = timed(lambda: tr.step(x), reps)
            row["step"] = {"ms": t_s, "patches_per_s": n / t_s * 1e3,
                           "sum_of_kernels_ms": ms + 2 * t_f + t_a + t_ad + t_n}
            if n <= 65536:                          # launch-bound shapes: CUDA-graph replay of the step
                trg = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9, use_cuda_graph="alias")
                t_g = timed(lambda: trg.step(x), max(reps, 50), warm=5)
                row["step_cuda_graph"] = {"ms": t_g, "patches_per_s": n / t_g * 1e3}
        else:
            idx = ops.bmu(x, geom, wd, cn)
            t_h = timed(lambda: ops.histogram(idx, k), reps)
            row["histogram"] = {"ms": t_h, "hbm_frac": (8.0 * n + 8.0 * k) / t_h / 1e6 / hbm}
        rows[name] = row
        print(name, json.dumps(row), flush=True)
        del x, cb, w, pool
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"peaks": {"hbm_gbs": hbm, "tc_3xtf32_tflops": tc_roof}, "configs": rows}, f, indent=1)


if __name__ == "__main__":
    main()
