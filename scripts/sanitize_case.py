"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck):
3D register-stencil kernels (TMA tiles and boundary tiles, 4x4x2 and dense 2x2x2 supercells, complex passes),
2D register-stencil, 1D, team kernels; plans, batch offsets, fastsum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nfft_b200 as T
torch.manual_seed(0)
dev = "cuda"
for d, N, m, B, n, C, cplx in [(3, 32, 4, 2, 3000, 1, False), (3, 8, 4, 1, 12000, 1, False), (3, 16, 3, 1, 800, 2, True),
                               (2, 32, 4, 2, 1500, 3, False), (1, 64, 8, 2, 500, 1, False), (3, 16, 6, 1, 500, 1, False)]:
    pos = torch.rand(n * B, d, device=dev) - 0.5
    batch = torch.arange(n * B, device=dev) // n
    ptr = torch.arange(B + 1, device=dev, dtype=torch.int64) * n
    x = torch.randn(n * B, C, device=dev, dtype=torch.complex64 if cplx else torch.float32)
    plan = T.NfftPlan(pos, batch_ptr=ptr)
    y = T.nfft_adjoint(x, plan=plan, N=N, m=m)
    f = T.nfft_forward(y, plan=plan, m=m, real_output=not cplx)
    y2 = T.nfft_adjoint(x, pos, batch, N, m)
    co = T.gaussian_analytic_coeffs(0.1, d, N)
    s = T.nfft_fastsum(x, co, pos * 0.5, batch=batch, cutoff=m)
    torch.cuda.synchronize()
    print(d, N, m, "ok", float(y.abs().sum()), float((y - y2).abs().max()), float(f.abs().sum()), float(s.abs().sum()),
          "dropped", plan.dropped_points(), flush=True)
