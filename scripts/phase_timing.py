"""Phase breakdown of the register-stencil kernels (debug build with -DNFFT_PHASE_TIMING).

Usage: NFFTB200_LIB=<debug .so> python scripts/phase_timing.py [c4|c4_clustered]
Prints the average clock64() cycles a CTA spends in each phase (see window_reg.cuh).
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nfft_b200 as T
from torch_nfft_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
d, N, m, n, B, C = 3, 128, 4, 2 ** 24, 4, 1
gen = torch.Generator(device="cuda"); gen.manual_seed(0)
if name.endswith("clustered"):
    K = 64
    centers = torch.rand(K, d, device="cuda", generator=gen) * 0.8 - 0.4
    ids = torch.randint(0, K, (n,), device="cuda", generator=gen)
    pos = ((centers[ids] + 0.02 * torch.randn(n, d, device="cuda", generator=gen) + 0.5) % 1.0) - 0.5
else:
    pos = torch.rand(n, d, device="cuda", generator=gen) - 0.5
x = torch.randn(n, C, device="cuda", generator=gen)
batch = (torch.arange(n, device="cuda") // (n // B)).contiguous()
lib = _lib.lib()
fn = lib.nfftb200_debug_phase_read
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 48)()
for it in range(3):
    y = T.nfft_adjoint(x, pos, batch, N, m, batch_size=B)
    z = T.nfft_forward(y, pos, batch, m, real_output=True, batch_size=B)
    torch.cuda.synchronize()
    fn(buf, 1)
for k, kern in enumerate(("spread", "gather")):
    v = [buf[k * 24 + i] for i in range(24)]
    ctas = max(v[4], 1)
    tot = sum(v[:4])
    print(kern, "CTAs", v[4], "cycles/CTA: setup %.0f sweep %.0f wait %.0f flush %.0f total %.0f" % (
        v[0] / ctas, v[1] / ctas, v[2] / ctas, v[3] / ctas, tot / ctas),
        "shares: " + " ".join("%.1f%%" % (100.0 * a / max(tot, 1)) for a in v[:4]),
        "| sweep length by warp: " + " ".join("%.0f" % (a / ctas) for a in v[8:16]),
        "| longest CTA %d cycles, kernel span %.3f ms, slot occupancy %.1f%% (2 CTAs x 148 SMs at 1.965 GHz), CTAs by fill quartile %s" % (
            v[16], (v[18] - ((~v[17]) & (2 ** 64 - 1))) * 1e-6,
            100.0 * tot / (296 * 1.965e9 * max((v[18] - ((~v[17]) & (2 ** 64 - 1))) * 1e-9, 1e-12)), v[19:23]),
        "| setup split: zero/tile-load %.0f bucket %.0f order %.0f" % (v[5] / ctas, v[6] / ctas, v[7] / ctas))
