"""BASELINE config c5: 3D fastsum (Gaussian kernel), N=64, m=4, n=2^26 points, point-sharded over the
ranks of one box (torchrun) with one NCCL all-reduce of the 128^3 grid per product; strong scaling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch_nfft_b200 as T
from torch_nfft_b200 import dist as D

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_total = 2 ** int(os.environ.get("C5_LOG2N", 26))
n = n_total // world
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
pos = (torch.rand(n, 3, device=dev, generator=g) - 0.5) * 0.5      # scaled into [-1/4, 1/4]^3
x = torch.randn(n, 1, device=dev, generator=g)
coeffs = T.gaussian_interpolated_coeffs(0.1, 3, 64)

def step():
    return D.nfft_fastsum_point_sharded(x, coeffs, pos, cutoff=4, batch_size=1)

from torch_nfft_b200 import _lib
for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_read()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 5
e0.record()
for _ in range(K):
    y = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
prof = _lib.profile_read()
if rank == 0:
    print("   stages (ms per product):", {k: round(v[0] / K, 3) for k, v in prof.items() if v[1]})
    print(f"c5 fastsum n=2^{n_total.bit_length()-1} on {world} GPU(s): {ms.item():.3f} ms per product, "
          f"{n_total / (ms.item() * 1e-3):.3e} points/s")
if world > 1:
    dist.destroy_process_group()
