"""A few fused SOM training steps of one shape: the command ncu wraps for launch lists.
usage: python tools/step_once.py <fmaps> <P> <K> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantized-autoregression-image-generator_b200")]
import torch  # noqa: E402
import somcb  # noqa: E402

b, p, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
d = 4 * p * p
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.tanh(torch.randn(b, 4, 32, 32, generator=g, device="cuda"))
pool = torch.tanh(torch.randn(max(8, k * d // 4096 + 1), 4, 32, 32, generator=g, device="cuda"))
cb = somcb.Codebook(patch_dim=(p, p), image_dim=(32, 32), image_channel=4, num_embeddings=k,
                    init_neighbour_range=k // 2).cuda()
with torch.no_grad():
    cb.codebook.weight.copy_(somcb.patchify(pool, (p, p)).reshape(-1, d)[:k])
tr = somcb.SomTrainer(cb, lr=1e-4, neighbourhood_step=10 ** 9)
for _ in range(steps):
    loss = tr.step(x)
torch.cuda.synchronize()
print("ok", float(loss))
