"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nfft_b200 as T
torch.manual_seed(0)
dev = "cuda"
for d, N, m, B, n, C, cplx in [(3, 32, 4, 2, 3000, 1, False), (3, 16, 3, 1, 800, 2, True), (2, 32, 4, 2, 1500, 3, False), (1, 64, 8, 2, 500, 1, False)]:
    pos = torch.rand(n * B, d, device=dev) - 0.5
    batch = torch.arange(n * B, device=dev) // n
    x = torch.randn(n * B, C, device=dev, dtype=torch.complex64 if cplx else torch.float32)
    y = T.nfft_adjoint(x, pos, batch, N, m)
    f = T.nfft_forward(y, pos, batch, m, real_output=not cplx)
    co = T.gaussian_analytic_coeffs(0.1, d, N)
    s = T.nfft_fastsum(x, co, pos * 0.5, batch=batch, cutoff=m)
    torch.cuda.synchronize()
    print(d, N, m, "ok", float(y.abs().sum()), float(f.abs().sum()), float(s.abs().sum()), flush=True)
