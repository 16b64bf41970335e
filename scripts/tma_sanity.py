"""First contact of a new kernel build with the GPU: small 3D transforms through the TMA plane flush / plane load of
the register-stencil kernels, checked against the oracle, before anything bigger runs (run under `timeout`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nfft_oracle as O  # noqa: E402
import torch_nfft_b200 as T  # noqa: E402

rng = np.random.default_rng(0)
worst = 0.0
for N, m, B, n in [(32, 4, 1, 3000), (32, 4, 2, 4000), (64, 4, 1, 20000), (16, 3, 2, 600), (32, 2, 1, 2000), (8, 4, 1, 40000)]:
    pos = rng.random((n * B, 3), dtype=np.float32) - 0.5
    pos[::5] = (pos[::5] * 0.1 + 0.3).astype(np.float32)          # a dense clump: heavy tiles, several chunks
    batch = np.repeat(np.arange(B, dtype=np.int64), n)
    x = rng.standard_normal((n * B, 1)).astype(np.float32)
    tp, tb, tx = (torch.from_numpy(a).cuda() for a in (pos, batch, x))
    plan = T.NfftPlan(tp, tb)
    y = T.nfft_adjoint(tx, plan=plan, N=N, m=m)
    f = T.nfft_forward(y, plan=plan, m=m, real_output=True)
    torch.cuda.synchronize()
    ry = O.nfft_adjoint(x, pos, batch, N, m)
    e1 = O.rel_l2(y.cpu().numpy(), ry)
    e2 = O.rel_l2(f.cpu().numpy(), O.nfft_forward(ry, pos, batch, m, real_output=True))
    worst = max(worst, e1, e2)
    print(f"N={N} m={m} B={B} n={n}: adjoint {e1:.2e} forward {e2:.2e} dropped {plan.dropped_points()}", flush=True)
print("tma_sanity", "OK" if worst < 1e-5 else "FAILED", worst)
sys.exit(0 if worst < 1e-5 else 1)
