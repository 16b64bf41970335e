"""Relative L2 error of the engine against the fp64 oracle and the fp32 oracle on a mid-size 3D and 2D
case (development aid for numerics experiments: NFFTB200_LIB selects the library variant)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_nfft_b200 as T
from oracle import nfft_oracle as O

rng = np.random.default_rng(3)
for d, N, m, n, C in ((3, 32, 4, 20000, 1), (3, 32, 3, 20000, 1), (2, 64, 4, 20000, 4)):
    pos = (rng.random((n, d)) - 0.5).astype(np.float32)
    x = rng.standard_normal((n, C)).astype(np.float32)
    y = T.nfft_adjoint(torch.from_numpy(x).cuda(), torch.from_numpy(pos).cuda(), None, N, m).cpu().numpy()
    y32 = O.nfft_adjoint(x, pos, None, N, m)
    y64 = O.nfft_adjoint(x, pos, None, N, m, prec="f64")
    f = T.nfft_forward(torch.from_numpy(y32).cuda(), torch.from_numpy(pos).cuda(), None, m).cpu().numpy()
    f32 = O.nfft_forward(y32, pos, None, m)
    f64 = O.nfft_forward(y32, pos, None, m, prec="f64")
    print(f"d={d} N={N} m={m}: adjoint vs f32 oracle {O.rel_l2(y, y32):.2e} vs f64 {O.rel_l2(y, y64):.2e} "
          f"(f32 oracle vs f64 {O.rel_l2(y32, y64):.2e}) | forward vs f32 {O.rel_l2(f, f32):.2e} vs f64 {O.rel_l2(f, f64):.2e} "
          f"(f32 vs f64 {O.rel_l2(f32, f64):.2e})")
