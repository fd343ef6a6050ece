"""Achieved HBM bandwidth of the HBM-bound kernels of the SOM path on ONE B200, at working sets larger than
the 126 MB L2 (so a repeat of the same launch cannot be served from cache).

    python tools/bench_hbm.py [--out gpurun_out/hbm.json] [--only adam,hist]

Each row: algorithmic bytes per launch (SURVEY 8d / DESIGN 5 per-unit figure x units), CUDA-event time per
launch after 3 warm-ups, GB/s and the fraction of MEASURED_PEAKS.json's HBM copy rate.
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


def timed(fn, reps=10, warm=3):
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "hbm.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        hbm = json.load(f)["hbm_gbs"]
    dev = torch.device("cuda", 0)
    rows = {}

    def want(name):
        return only is None or name in only

    def record(name, nbytes, ms, **kw):
        rows[name] = dict(bytes=nbytes, ms=ms, gbs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / hbm, **kw)
        print(name, json.dumps(rows[name]), flush=True)

    g = torch.Generator(device=dev).manual_seed(5)

    # reference point: a plain device-to-device copy of 512 MB (read + write)
    if want("copy"):
        a = torch.empty(1 << 27, device=dev)
        b = torch.empty_like(a)
        record("copy_512MB", 2 * a.numel() * 4, timed(lambda: b.copy_(a)))
        # write-only and read-only reference points (torch fill / memset / sum), to tell a write-bound kernel from a
        # slow one: HBM write streams do not reach the mixed copy rate
        record("fill_512MB_write_only", a.numel() * 4, timed(lambda: a.fill_(1.0)))
        record("memset_512MB_write_only", a.numel() * 4, timed(lambda: a.zero_()))
        record("sum_512MB_read_only", a.numel() * 4, timed(lambda: a.sum()))
        del a, b

    if want("adam"):
        n = 1 << 26
        w, m, v = (torch.randn(n, device=dev, generator=g) * 0.01 for _ in range(3))
        v.abs_()
        gr = torch.randn(n, device=dev, generator=g)
        record("adam_64M", 28 * n, timed(lambda: ops.adam_step(w, m, v, gr, 1e-4, 3)))
        del w, m, v, gr

    if want("norm2"):
        for k, d in ((262144, 256), (1 << 22, 16), (1 << 20, 64), (16384, 4096)):
            w = torch.randn(k, d, device=dev, generator=g)
            record("norm2_K%d_D%d" % (k, d), 4 * k * d + 4 * k, timed(lambda: ops.prepare_codebook(w)))
            del w

    if want("merge"):
        r, n = 8, 1 << 22
        rd = torch.rand(r, n, device=dev, generator=g)
        idx = torch.randint(0, 262144, (r, n), device=dev, generator=g)
        record("merge_R8_4M", 12 * r * n + 12 * n, timed(lambda: ops.merge_candidates(rd, idx)))
        del rd, idx

    if want("hist"):
        n = 1 << 26
        for k in (1024, 4096, 32768, 262144):
            idx = torch.randint(0, k, (n,), device=dev, generator=g)
            cnt = torch.zeros(k, dtype=torch.int64, device=dev)
            record("hist_uniform_K%d_64M" % k, 8 * n + 8 * k, timed(lambda: ops.histogram(idx, k, cnt)))
            # skewed hits: squared uniform concentrates on the low units (a trained map's dense region)
            sk = (torch.rand(n, device=dev, generator=g).pow_(3) * k).long().clamp_(0, k - 1)
            record("hist_skewed_K%d_64M" % k, 8 * n + 8 * k, timed(lambda: ops.histogram(sk, k, cnt)))
            del idx, sk, cnt
        idx = torch.randint(0, 32768, (1 << 20,), device=dev, generator=g)
        cnt = torch.zeros(32768, dtype=torch.int64, device=dev)
        record("hist_C5_1M_K32768", 8 * (1 << 20) + 8 * 32768, timed(lambda: ops.histogram(idx, 32768, cnt)),
               note="L2-resident / launch-bound size")
        del idx, cnt

    if want("quantize"):
        for name, fm, pd, k in (("C4", 16384, (4, 4), 16384), ("C2", 39063, (2, 2), 4096),
                                ("C5", 16384, (8, 8), 32768), ("C3", 4096, (32, 32), 512)):
            d = 4 * pd[0] * pd[1]
            geom = ops.geometry((fm, 4, 32, 32), pd)
            n = ops.n_patches_of(geom)
            table = torch.randn(k, d, device=dev, generator=g)
            idx = torch.randint(0, k, (n,), device=dev, generator=g)
            out = torch.empty(fm, 4, 32, 32, device=dev)
            record("quantize_%s" % name, (4 * d + 8) * n, timed(lambda: ops.quantize(idx, table, geom, out=out)),
                   note="write 4D + read 8 per patch; table rows from L2")
            del table, idx, out

    if want("gather"):
        k, d = 262144, 256
        w = torch.randn(k, d, device=dev, generator=g)
        keep = torch.nonzero(torch.rand(k, device=dev, generator=g) < 0.5).flatten()
        record("gather_rows_K262144_D256", 8 * d * keep.numel() + 8 * keep.numel(),
               timed(lambda: ops.gather_rows(w, keep)))
        del w, keep

    if want("tokens"):
        n, ls, hs = 1 << 16, 64, 256
        lr = torch.randint(0, 1024, (n, ls), device=dev, generator=g)
        hr = torch.randint(0, 4096, (n, hs), device=dev, generator=g)
        nbytes = 8 * n * (ls + hs) + 8 * n * (ls + hs + hs + 1)
        record("assemble_tokens_base_64K", nbytes, timed(lambda: ops.assemble_tokens(lr, hr, 1024, 4096, True)))
        del lr, hr

    if want("accumulate"):
        for name, fm, pd, k in (("C4", 16384, (4, 4), 16384), ("C2", 39063, (2, 2), 4096),
                                ("C5", 16384, (8, 8), 32768)):
            d = 4 * pd[0] * pd[1]
            x = torch.tanh(torch.randn(fm, 4, 32, 32, device=dev, generator=g))
            geom = ops.geometry(x.shape, pd)
            n = ops.n_patches_of(geom)
            table = torch.randn(k, d, device=dev, generator=g)
            idx = torch.randint(0, k, (n,), device=dev, generator=g)
            ms = timed(lambda: ops.accumulate(x, geom, idx, table, k, want_sse=True))
            record("accumulate_%s" % name, (4 * d + 8) * n + 8 * k * d, ms)
            del x, table, idx

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"hbm_gbs_peak": hbm, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
