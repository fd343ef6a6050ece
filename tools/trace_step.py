"""Device-timestamp trace of the C4 training step (any number of ranks): where the time of a replayed step graph goes,
kernel by kernel, on every rank -- CUDA events cannot time inside a graph and ncu cannot profile a multi-rank run.
usage: [torchrun ...] python tools/trace_step.py [tail] [wt]      (tail: auto | nccl | peer, wt: full | slice)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import somcb  # noqa: E402
import bench  # noqa: E402

NAMES = {1: "filter: split_in_t", 2: "filter: tc main", 3: "filter: reduce", 4: "filter: ffma", 5: "norms",
         6: "bmu: operand split", 7: "bmu: main", 8: "acc: pairs / histogram (+ radix sort when K > 16384)", 9: "acc: offsets",
         10: "acc: level 1", 11: "acc: level 2", 12: "acc: sse reduce", 13: "peer: barrier start",
         113: "peer: barrier passed", 14: "peer: reduce rows", 15: "peer: adam slice + multicast + barrier",
         16: "peer: multicast rows + barrier", 17: "peer: all-reduce", 18: "adam (dp)", 19: "adam", 20: "acc: sort prefix", 21: "acc: sort offsets", 22: "acc: sort scatter",
         121: "acc: (end of sort offsets)", 122: "acc: (end of sort scatter, block 0)"}


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tail = sys.argv[1] if len(sys.argv) > 1 else "auto"
    wt = sys.argv[2] if len(sys.argv) > 2 else "full"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = ctypes.CDLL(somcb._lib.LIB_PATH)
    if os.environ.get("TRACE_PDL", "1") == "0":            # A/B: plain stream order between the step's kernels
        lib.som_debug_set_pdl(0)
    lo, hi = somcb.shard_bounds(int(os.environ.get("TRACE_IMAGES", "16384")), world, rank)
    xs = [bench._fmaps(hi - lo, 5000 + 131 * b + rank, dev) for b in range(4)]
    cb = bench._codebook(16384, (4, 4), dev)
    tr = bench._make_trainer(cb, world, tail=tail, wt=wt)
    if os.environ.get("TRACE_OVERLAP", "1") == "0":        # A/B: W~ filter in front of the search, on one stream
        tr.overlap_filter = False
    if world > 1:
        tr.broadcast_weights(0)
    for i in range(10):
        tr.step(xs[i % 4])
    torch.cuda.synchronize()
    buf = torch.zeros(8200, dtype=torch.int64, device=dev)
    lib.som_debug_trace(ctypes.c_void_p(buf.data_ptr()))
    steps = 5
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        tr.step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    lib.som_debug_trace(ctypes.c_void_p(0))
    b = buf.cpu().tolist()
    n = b[0]
    ev = [(b[1 + 2 * i], b[2 + 2 * i]) for i in range(n)]
    ev.sort(key=lambda t: t[1])
    per = n // steps
    rows = {}
    order = []
    for s in range(1, steps):                        # skip the first traced step (its predecessor's tail is missing)
        seg = ev[s * per:(s + 1) * per]
        nxt = ev[(s + 1) * per][1] if (s + 1) * per < n else None
        for i, (kid, t) in enumerate(seg):
            t_next = seg[i + 1][1] if i + 1 < len(seg) else nxt
            if t_next is None:
                continue
            key = (i, kid)
            if key not in rows:
                rows[key] = []
                order.append(key)
            rows[key].append((t_next - t) / 1e3)
    out = {"rank": rank, "world": world, "tail": getattr(tr, "tail", "single"), "wt": wt,
           "step_ms_events": e0.elapsed_time(e1) / steps,
           "kernels_us": [[NAMES.get(k[1], str(k[1])), round(sum(v) / len(v), 1)] for k, v in ((k, rows[k]) for k in order)]}
    out["sum_us"] = round(sum(v for _, v in out["kernels_us"]), 1)
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        if world > 1:
            dist.barrier()
    if world > 1:
        tr._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
