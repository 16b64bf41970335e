"""Multi-GPU check of torch_nfft_b200.dist (run with torchrun, one rank per GPU):
point-sharded adjoint / forward / fastsum and batch-sharded adjoint against the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import torch_nfft_b200 as T
from torch_nfft_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev); g.manual_seed(0)          # same inputs on every rank
d, N, m, B, n = 3, 64, 4, 2, 400000
pos = torch.rand(n, d, device=dev, generator=g) - 0.5
x = torch.randn(n, 1, device=dev, generator=g)
batch = torch.arange(n, device=dev) // (n // B)
rel = lambda a, b: (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()
full = T.nfft_adjoint(x, pos, batch, N, m, batch_size=B)
lo, hi = D.shard_points(n, world, rank)
y = D.nfft_adjoint_point_sharded(x[lo:hi], pos[lo:hi], batch[lo:hi], N, m, batch_size=B)
e1 = rel(y, full)
f_full = T.nfft_forward(full, pos, batch, m, real_output=True, batch_size=B)
f = D.nfft_forward_point_sharded(full, pos[lo:hi], batch[lo:hi], m, real_output=True)
e2 = rel(f, f_full[lo:hi])
co = T.gaussian_analytic_coeffs(0.1, d, N)
s_full = T.nfft_fastsum(x, co, pos * 0.5, batch=batch, cutoff=m, batch_size=B)
s = D.nfft_fastsum_point_sharded(x[lo:hi], co, pos[lo:hi] * 0.5, source_batch=batch[lo:hi], cutoff=m, batch_size=B)
e3 = rel(s, s_full[lo:hi])
yb, (b_lo, b_hi) = D.nfft_adjoint_batch_sharded(x, pos, batch, N, m, batch_size=B, gather_output=True)
e4 = rel(yb, full)
# B a multiple of the world size: reduce-scatter into whole grids per rank, spectra all-gathered / kept local
Bw = world
nw = (n // Bw) * Bw
batch_w = torch.arange(nw, device=dev) // (nw // Bw)
full_w = T.nfft_adjoint(x[:nw], pos[:nw], batch_w, N, m, batch_size=Bw)
lo, hi = D.shard_points(nw, world, rank)
yw = D.nfft_adjoint_point_sharded(x[lo:hi], pos[lo:hi], batch_w[lo:hi], N, m, batch_size=Bw)
e5 = rel(yw, full_w)
ywl, (o_lo, o_hi) = D.nfft_adjoint_point_sharded(x[lo:hi], pos[lo:hi], batch_w[lo:hi], N, m, batch_size=Bw,
                                                 scatter_output=True)
e6 = rel(ywl, full_w[o_lo:o_hi])
errs = torch.tensor([e1, e2, e3, e4, e5, e6], device=dev)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    print("dist check (max over ranks) adjoint_point %.2e forward_point %.2e fastsum_point %.2e adjoint_batch %.2e "
          "adjoint_point_reduce_scatter %.2e (local %.2e)" % tuple(errs.tolist()))
    assert errs.max().item() < 1e-5
dist.destroy_process_group()
