"""Quick CUDA-event timing of adjoint / forward at the BASELINE configs (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_nfft_b200 as T

CONFIGS = {
    "c2": (1, 1024, 8, 2 ** 20, 64, 1),
    "c3": (2, 256, 4, 2 ** 23, 16, 8),
    "c4": (3, 128, 4, 2 ** 24, 4, 1),
    "c5gpu": (3, 64, 4, 2 ** 23, 1, 1),
    "c4small": (3, 128, 4, 2 ** 20, 4, 1),
    "c4m3": (3, 128, 3, 2 ** 24, 4, 1),   # the reference's default cutoff
    "c4m2": (3, 128, 2, 2 ** 24, 4, 1),
    "c3m3": (2, 256, 3, 2 ** 23, 16, 8),
    "c4m6": (3, 128, 6, 2 ** 22, 4, 1),   # m > 4: team kernels
}


def clustered(n, d, gen):
    K = 64
    centers = torch.rand(K, d, device="cuda", generator=gen) * 0.8 - 0.4
    ids = torch.randint(0, K, (n,), device="cuda", generator=gen)
    p = centers[ids] + 0.02 * torch.randn(n, d, device="cuda", generator=gen)
    return ((p + 0.5) % 1.0) - 0.5


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name in (sys.argv[1:] or ["c4small", "c4", "c3", "c2", "c5gpu"]):
    dist = "uniform"
    if name.endswith("_clustered"):
        name, dist = name[:-10], "clustered"
    d, N, m, n, B, C = CONFIGS[name]
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    pos = (torch.rand(n, d, device="cuda", generator=gen) - 0.5) if dist == "uniform" else clustered(n, d, gen)
    x = torch.randn(n, C, device="cuda", generator=gen)
    batch = (torch.arange(n, device="cuda") // (n // B)).contiguous()
    y = T.nfft_adjoint(x, pos, batch, N, m, batch_size=B)
    t_adj = timeit(lambda: T.nfft_adjoint(x, pos, batch, N, m, batch_size=B))
    t_fwd = timeit(lambda: T.nfft_forward(y, pos, batch, m, real_output=True, batch_size=B))
    def pair():
        yy = T.nfft_adjoint(x, pos, batch, N, m, batch_size=B)
        return T.nfft_forward(yy, pos, batch, m, real_output=True, batch_size=B)
    t_pair = timeit(pair)
    print(f"{name} {dist}: adjoint {t_adj:.3f} ms  forward {t_fwd:.3f} ms  pair(alternating) {t_pair:.3f} ms  "
          f"{n/(t_pair*1e-3):.3e} pts/s", flush=True)
    del x, pos, batch, y
    torch.cuda.empty_cache()
