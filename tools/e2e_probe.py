"""End-to-end tokenisation from pinned host memory: chunk-size sweep + the plain H2D copy bandwidth of the box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch
import somcb
dev = torch.device("cuda", 0)
n = 39063
host = torch.empty(n, 4, 32, 32, pin_memory=True).normal_().tanh_()
out = torch.empty(n, 256, dtype=torch.int64, pin_memory=True)
xd = torch.empty(n, 4, 32, 32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    xd.copy_(host, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    xd.copy_(host, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"plain H2D 640 MB: {ms:.3f} ms -> {host.numel() * 4 / ms / 1e6:.1f} GB/s")
cb = somcb.Codebook(patch_dim=(2, 2), image_dim=(32, 32), image_channel=4, num_embeddings=4096, init_neighbour_range=2048).to(dev)
with torch.no_grad():
    cb.codebook.weight.copy_(somcb.patchify(xd[:64], (2, 2)).reshape(-1, 16)[:4096])
for chunk, depth in ((4096, 3), (2048, 3), (2048, 4), (1024, 4), (8192, 3)):
    tok = somcb.HostTokenizer(cb, chunk_fmaps=chunk, depth=depth)
    for _ in range(2):
        tok.tokenize(host, out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(8):
        tok.tokenize(host, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    print(f"chunk {chunk} depth {depth}: {ms:.3f} ms -> {n * 256 / ms / 1e6:.3f} G patches/s")
