#!/usr/bin/env python
"""bench.py -- SOM-codebook hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Workload at every N (weak scaling, patches shard across ranks, no data-path collective):
BASELINE.json configs[1] -- fine-patch tokenisation, P=2 (D=16), 4096-unit codebook, 39 063
synthetic 4x32x32 latent fmaps = 10 000 128 patches per GPU per step, BMU only.  One step = one pass
of Codebook.get_patches_bmu over that batch.  `value` has the inputs resident in HBM; `e2e` runs
the same batch from pinned HOST memory through somcb.HostTokenizer (H2D + BMU + D2H of int64
indices inside the timed region).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "quantized-autoregression-image-generator_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# NCCL prints "NCCL version ..." on STDOUT at NCCL_DEBUG=VERSION (and INFO); the contract is ONE JSON line there.
# An explicit INFO / TRACE request is left alone.
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

import torch  # noqa: E402

METRIC = "patches/sec BMU (fine-patch tokenisation P=2 D=16 K=4096, 10M patches/step/GPU)"
UNIT = "patches/s"
C2 = dict(n_fmaps=39063, patch=(2, 2), K=4096, C=4, H=32, W=32)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # derived: 74.4 at max clock


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        names = {}
        for nm in dir(nv):
            if nm.startswith("nvmlClocksEventReason") or nm.startswith("nvmlClocksThrottleReason"):
                val = getattr(nv, nm)
                if isinstance(val, int) and val:
                    names.setdefault(val, nm.replace("nvmlClocksEventReason", "")
                                     .replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit and bit & (bit - 1) == 0:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(s)}


def _trained_like_codebook(k, patch, seed=7):
    """Seeded synthetic codebook for OUR arm: K distinct data patches drawn from an independent pool of
    tanh(randn) fmaps.  Same recipe as the oracle's helper, restated here so that the measured
    arm imports nothing from oracle/."""
    import somcb
    p_h, p_w = patch
    seq = (32 // p_h) * (32 // p_w)
    n_fmaps = max(2, (2 * k + seq - 1) // seq)
    g = torch.Generator().manual_seed(1000 + seed)
    pool = somcb.patchify(torch.tanh(torch.randn(n_fmaps, 4, 32, 32, generator=g)), patch)
    pool = pool.reshape(-1, pool.shape[-1])
    g2 = torch.Generator().manual_seed(int(seed))
    pick = torch.randperm(pool.shape[0], generator=g2)[:k]
    return pool[pick].clone().contiguous()


def _c2_inputs(dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.empty(C2["n_fmaps"], C2["C"], C2["H"], C2["W"], device=dev)
    for lo in range(0, C2["n_fmaps"], 8192):
        hi = min(C2["n_fmaps"], lo + 8192)
        x[lo:hi] = torch.tanh(torch.randn(hi - lo, C2["C"], C2["H"], C2["W"], generator=g, device=dev))
    return x


def _c2_codebook(dev=None):
    import somcb
    w = _trained_like_codebook(C2["K"], C2["patch"], 7)
    cb = somcb.Codebook(patch_dim=C2["patch"], image_dim=(C2["H"], C2["W"]), image_channel=C2["C"],
                        num_embeddings=C2["K"], init_neighbour_range=C2["K"] // 2)
    with torch.no_grad():
        cb.codebook.weight.copy_(w)
    return (cb.to(dev) if dev is not None else cb), w


def _cpu_reference_bmu(steps, warmup, sample_fmaps=256, min_seconds=0.0):
    """The reference's CPU path for this workload (oracle port, all host threads): patches/s."""
    from oracle.step_oracle import make_oracle_codebook, synthetic_fmaps, trained_like_codebook
    torch.set_num_threads(os.cpu_count() or 1)
    w = trained_like_codebook(C2["K"], C2["patch"], 7)
    cb = make_oracle_codebook(w, C2["patch"], (C2["H"], C2["W"]), C2["C"], C2["K"] // 2)
    x = synthetic_fmaps(sample_fmaps, 123)
    n_p = sample_fmaps * 256
    with torch.no_grad():
        for _ in range(warmup):
            cb.get_patches_bmu(x, reshape=True)
        t0 = time.perf_counter()
        done = 0
        while done < steps or (time.perf_counter() - t0) < min_seconds:
            cb.get_patches_bmu(x, reshape=True)
            done += 1
        dt = time.perf_counter() - t0
    return {"value": n_p * done / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{done} calls x {n_p} patches (of the 10 000 128-patch step), oracle restatement of "
                      f"models/Codebook.py:77-99 on torch {torch.__version__} CPU, {dt:.1f} s"}, dt / done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = _cpu_reference_bmu(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "BASELINE configs[1] (C2): P=2 D=16 K=4096 BMU-only, "
                                            "each step a 65 536-patch sample of the 10 000 128-patch batch"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _extra_training(dev, world, rank, group):
    """BMU + update step (BASELINE config 4 shape: P=4 D=64 K=16384, 2^20 patches per GPU per step)."""
    import somcb
    k, pd, n_f = 16384, (4, 4), 16384
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.tanh(torch.randn(n_f, 4, 32, 32, generator=g, device=dev))
    cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                        init_neighbour_range=k // 2)
    with torch.no_grad():
        cb.codebook.weight.copy_(_trained_like_codebook(k, pd, 7))
    cb = cb.to(dev)
    tr = somcb.DataParallelSom(cb, lr=1e-4, neighbourhood_step=200) if world > 1 else \
        somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=200)
    for _ in range(3):
        tr.step(x)
    steps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = tr.step(x)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms) / steps
    return {"workload": "C4 shape: SOM step (BMU + accumulate + 2 filters + Adam), P=4 D=64 K=16384, "
                        "1 048 576 patches/GPU/step" + (", all-reduce of Rbar per step" if world > 1 else ""),
            "patches_per_s": world * n_f * 64 / (ms * 1e-3), "ms_per_step": ms, "loss": float(loss)}


def _extra_configs(dev):
    """Single-GPU timings of the other BASELINE shapes (device time, CUDA events): C1 step with CUDA-graph
    replay, C3 BMU + full step, C5 one GPU's shard BMU.  Reported beside the headline, not part of it."""
    import somcb
    from somcb import ops

    def data(n, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        x = torch.empty(n, 4, 32, 32, device=dev)
        for lo in range(0, n, 8192):
            hi = min(n, lo + 8192)
            x[lo:hi] = torch.tanh(torch.randn(hi - lo, 4, 32, 32, generator=g, device=dev))
        return x

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

    def codebook(k, pd, rng):
        d = 4 * pd[0] * pd[1]
        pool = data(max(8, (k * d) // 4096 + 1), 7)
        w = somcb.patchify(pool, pd).reshape(-1, d)[:k].contiguous()
        cb = somcb.Codebook(patch_dim=pd, image_dim=(32, 32), image_channel=4, num_embeddings=k,
                            init_neighbour_range=rng).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(w)
        return cb

    out = {}
    x1, cb1 = data(8, 123), codebook(1024, (4, 4), 512)
    tr1 = somcb.SomTrainer(cb1, lr=1e-4, neighbourhood_step=10 ** 9, use_cuda_graph="alias")
    out["C1_step_cuda_graph_us"] = 1e3 * timed(lambda: tr1.step(x1), 100, warm=5)
    x3, cb3 = data(4096, 123), codebook(512, (32, 32), 256)
    w3 = cb3.codebook.weight.data
    g3 = ops.geometry(x3.shape, (32, 32))
    out["C3_bmu_ms"] = timed(lambda: ops.bmu(x3, g3, w3, ops.prepare_codebook(w3)), 20)
    tr3 = somcb.SomTrainer(cb3, lr=1e-4, neighbourhood_step=10 ** 9, use_cuda_graph="alias")
    out["C3_step_ms"] = timed(lambda: tr3.step(x3), 20)
    del x3, tr3, cb3
    x5, cb5 = data(65536, 123), codebook(32768, (8, 8), 16384)
    w5 = cb5.codebook.weight.data
    g5 = ops.geometry(x5.shape, (8, 8))
    cn5 = ops.prepare_codebook(w5)
    out["C5_shard_bmu_ms"] = timed(lambda: ops.bmu(x5, g5, w5, cn5), 3, warm=1)
    out["shapes"] = ("C1: 512 patches D=64 K=1024; C3: 4096 patches D=4096 K=512; "
                     "C5: 1 048 576 patches D=256, 32 768 of 262 144 units")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 FFMA, 2 tcgen05 3xTF32")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 10)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: somcb has no CPU fallback")
    import somcb
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = somcb._lib.load()
    peaks = _peaks()

    cb, w_cpu = _c2_codebook(dev)
    cb.bmu_variant = args.variant
    cb.eval()
    x = _c2_inputs(dev, 123 + rank)
    n_p = C2["n_fmaps"] * 256
    d_dim, k = 16, C2["K"]
    variant = args.variant or lib.som_bmu_pick_variant(n_p, d_dim, k)

    with torch.no_grad():
        for _ in range(args.warmup):
            idx = cb.get_patches_bmu(x, reshape=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = lib.som_launch_count()
        with ClockSampler(local) as clk:
            e0.record()
            for _ in range(args.steps):
                idx = cb.get_patches_bmu(x, reshape=True)
            e1.record()
            torch.cuda.synchronize()
        launches = lib.som_launch_count() - l0
        if world > 1:
            dist.barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t) / args.steps
    value = world * n_p / (ms_step * 1e-3)

    # sanity inside the bench: indices in range and the histogram adds up
    counts = somcb.ops.histogram(idx.reshape(-1), k)
    assert int(counts.sum()) == n_p

    # ---- end to end from pinned host memory through the public host API ----------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    # one process per GPU: stage out of the GPU's own NUMA node (no-op when the topology is not exposed)
    cpus_before = os.sched_getaffinity(0)
    numa_node = somcb.bind_host_to_gpu_node(dev)
    host = torch.empty(C2["n_fmaps"], C2["C"], C2["H"], C2["W"], pin_memory=True)
    host.copy_(x)
    out_host = torch.empty(C2["n_fmaps"], 256, dtype=torch.int64, pin_memory=True)
    tok = somcb.HostTokenizer(cb, chunk_fmaps=4096, depth=3)
    for _ in range(2):
        tok.tokenize(host, out_host)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        tok.tokenize(host, out_host)
    f1.record()
    torch.cuda.synchronize()
    assert torch.equal(out_host, idx.cpu()), "e2e indices differ from the resident-input run"
    os.sched_setaffinity(0, cpus_before)            # the CPU baseline below uses every host core again
    e2e_t = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_t) / e2e_steps
    e2e = {"value": world * n_p / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 8,
           "api": "somcb.HostTokenizer.tokenize(pinned fmaps) -> pinned int64 indices",
           "host_numa_node": numa_node}

    # ---- roofline of the dominant kernel (BMU) --------------------------------------------------
    flops = 2.0 * k * d_dim * n_p
    achieved = flops / (ms_step * 1e-3) / 1e12
    # config-S split mode of the library (csrc/som_bmu_tc_s.cu): FP16 hi/lo from 65 536 patches on, TF32 below or with SOM_TC_S_F16=0
    f16_split = variant == 2 and os.environ.get("SOM_TC_S_F16", "1") != "0" and n_p >= 65536
    bf16_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    tc_peak = bf16_peak / 3.0 if f16_split else bf16_peak * 0.5 / 3.0
    traffic, pipe_pct = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("bmu_c2_f16_dram_bytes_per_launch" if f16_split else "bmu_c2_dram_bytes_per_launch")
        pipe_pct = tj.get("bmu_c2_f16_tensor_pipe_active_pct" if f16_split else "bmu_c2_tensor_pipe_active_pct")
    # nominal fp32-faithful roof of this shape: 2048 TF32 MAC/clk/SM x 148 SMs x max clock / 3 products,
    # times the useful fraction of the K' = 3*16 + 8 inner dimension
    nominal = 2048 * 2 * 148 * 1.965e9 / 3.0 * (48.0 / 56.0) / 1e12
    # the tensor pipe's own sustained TF32 rate on this pool (tools/mma_rate.cu, power-capped clock), / 3 products
    pipe_peak = None
    ppath = os.path.join(ROOT, "profiles", "measured_pipe_peaks.json")
    if os.path.exists(ppath):
        with open(ppath) as f:
            pipe_peak = json.load(f).get("tf32_3x_fp32_faithful_tflops_sustained")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                "frac": achieved / tc_peak, "traffic": traffic,
                "kernel": ("bmu_tc_s (tcgen05 kind::f16 x3, FP16 hi/lo split)" if f16_split else
                           "bmu_tc3x (tcgen05 kind::tf32 x3)") if variant == 2 else "bmu_ffma (fp32 FFMA)",
                "peak_basis": (f"{peaks['_source']}: sustained bf16 x 1/3 (three 16-bit products, fp32-faithful)"
                               if f16_split else
                               f"{peaks['_source']}: sustained bf16 x 1/2 (tf32) x 1/3 (3xTF32, fp32-faithful)"),
                "algorithmic_flops_per_patch": 2 * k * d_dim,
                "algorithmic_bytes_per_patch": 4 * d_dim + 8,
                "hbm_frac": (4 * d_dim + 8) * n_p / (ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "ffma_frac_of_derived_74.4TF": achieved / FFMA_PEAK_TFLOPS,
                "frac_of_nominal_tf32_pipe": achieved / nominal,
                "frac_of_measured_tcgen05_tf32_peak": (achieved / pipe_peak) if pipe_peak else None,
                "measured_tcgen05_tf32_peak_3x": pipe_peak,
                "tensor_pipe_active_pct_ncu": pipe_pct,
                "note": ("FP16-split mode: 4 MMAs (512 tensor-pipe cycles) per 128x256 tile; the pace is set by the "
                         "epilogue's min-reduction on the half-rate ALU pipe (~850 cycles per tile, 547 with the "
                         "reduction switched off: DESIGN 5), so frac is against a roof this kernel does not bind on; "
                         "SOM_TC_S_F16=0 selects the 3xTF32 mode (7 MMAs per tile, tensor pipe 94% active)"
                         if f16_split else
                         "frac > 1 is expected: the denominator is the measured sustained cuBLAS bf16 rate / 6; "
                         "the kernel keeps the tensor pipe ~94% active (ncu) and is bounded by the SM clock "
                         "under the power cap (1.55-1.65 GHz).  Against the tensor pipe's own measured sustained "
                         "TF32 rate (930 TFLOP/s dense = 310 fp32-faithful) see frac_of_measured_tcgen05_tf32_peak; "
                         "7 MMAs per tile carry 6 MMAs of useful products (K' = 56 for 3*16)")}

    extra = None
    if not args.no_extra:
        try:
            del host, out_host, tok
            extra = _extra_training(dev, world, rank, None)
        except Exception as e:  # noqa: BLE001
            extra = {"error": repr(e)}

    if extra is not None and world == 1 and not args.no_extra:
        try:
            torch.cuda.empty_cache()
            extra["other_configs"] = _extra_configs(dev)
        except Exception as e:  # noqa: BLE001
            extra["other_configs"] = {"error": repr(e)}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = _cpu_reference_bmu(steps=4, warmup=1, min_seconds=10.0)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE configs[1] (C2): 39 063 synthetic 4x32x32 fmaps, P=2 (D=16), "
                                       "K=4096 trained-like codebook, 10 000 128 patches per GPU per step, BMU only",
                           "l2": "per-step input 640 MB + 80 MB of indices exceed the 126 MB L2",
                           "variant": int(variant), "patches_per_gpu": n_p},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_base, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
