"""Probe torch symmetric memory (peer pointers over NVLink, NVSwitch multicast) on this box.
Launch with torch.distributed.run; prints what the fused data-parallel tail can rely on."""
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << 20
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    info = {"rank": hdl.rank, "world": hdl.world_size, "buffer_ptrs": [hex(p) for p in hdl.buffer_ptrs],
            "signal_pad_size": hdl.signal_pad_size, "multicast": bool(hdl.has_multicast_support),
            "multicast_ptr": hex(hdl.multicast_ptr) if hdl.has_multicast_support else None}
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (n,), torch.float32)
    torch.cuda.synchronize()
    ok = bool((remote == float(peer + 1)).all())
    # peer read bandwidth: sum over all peers' buffers
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = torch.zeros(n, device=dev)
    hdl.barrier(channel=0)
    e0.record()
    for _ in range(20):
        for p in range(world):
            acc += hdl.get_buffer(p, (n,), torch.float32)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    info.update(peer_read_ok=ok, sum_over_peers_ms=ms, gbs_in=world * n * 4 / ms / 1e6)
    # NCCL all-reduce of the same 4 MB for comparison
    x = torch.ones(n + 4, device=dev)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    info["nccl_allreduce_4MB_us"] = e0.elapsed_time(e1) / 50 * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(10):
            dist.all_reduce(x)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    info["nccl_allreduce_4MB_graph_us"] = e0.elapsed_time(e1) / 50 * 1e3
    for r in range(world):
        if r == rank:
            print(info, flush=True)
        dist.barrier()
    del g
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
