#!/usr/bin/env python
"""Benchmark of the NFFT hot path (BASELINE.json: NU points/sec, adjoint + forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c4|c4_clustered|c3|c2] [--ref-device cuda|cpu]

One "step" = one adjoint NFFT (real x -> complex spectrum) followed by one forward NFFT
(spectrum -> real values at the same points) of the workload, i.e. pts/s = n / (t_adj + t_fwd).
Default workload (N = 1 and per GPU for N > 1): BASELINE configs[3] "c4": 3D, N=128, m=4,
n=2^24 uniform points, batch_size=4, 1 channel -- the configuration the north-star target is
quoted on.  For N > 1 every rank transforms its own 4 batch entries (batch sharding, no
data-path collective): weak scaling.

Printed JSON (rank 0, one line): see the keys below.  `value` has inputs resident in HBM;
`e2e` goes through the public API with pinned HOST buffers (H2D of pos / x / batch offsets and D2H of
both results inside the timed region).  `roofline` is for the dominant kernel (spread), from CUDA
events recorded inside the library around that stage during the timed region.

Every step bins its points once (`NfftPlan`, made inside the step) and uses that binning for the
adjoint and the forward transform of the pair; nothing is carried from one step to the next.  The
point sets are described by batch_size + 1 offsets (`batch_ptr`) instead of the reference's per-point
int64 vector, which the engine accepts too.

Extra objects in the same line: `extra_workloads` (N = 1: c4 Gaussian-clustered, c3, c2 eager and as a
CUDA graph), and for --gpus N > 1 `strong_c4` (the SAME 2^24-point c4 spread over the N GPUs: batch
entries first, then a 2-way point split per entry with a pairwise all-reduce of the 67 MB grid) and
`c5_point_sharded` (BASELINE config c5: 2^26-point 3D fastsum, points split over the N GPUs, one
all-reduce of the 8 MB grid per product), each with its own CUDA-event time and the collective's.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (d, N, m, n_total, B, C, distribution)
    "c2": (1, 1024, 8, 2 ** 20, 64, 1, "uniform"),
    "c3": (2, 256, 4, 2 ** 23, 16, 8, "uniform"),
    "c4": (3, 128, 4, 2 ** 24, 4, 1, "uniform"),
    "c4_clustered": (3, 128, 4, 2 ** 24, 4, 1, "clustered"),
    "c4_small": (3, 128, 4, 2 ** 20, 4, 1, "uniform"),
}


def algorithmic_bytes(d, N, n, B, C):
    """compulsory HBM traffic with real-data FFTs."""
    M = 2 * N
    G_r = B * C * M ** d * 4
    G_h = B * C * M ** (d - 1) * (M // 2 + 1) * 8
    Y_c = B * C * N ** d * 8
    Y_h = B * C * N ** (d - 1) * (N // 2 + 1) * 8
    adj = n * (4 * d + 4 * C + 8) + G_r + (G_r + G_h) + (Y_h + Y_c)
    fwd = (Y_c + G_h) + (G_h + G_r) + G_r + n * (4 * d + 8) + 4 * n * C
    spread = n * (4 * d + 4 * C + 8) + G_r       # pos + x + batch read, real grid written
    gather = G_r + n * (4 * d + 8) + 4 * n * C   # grid read, pos + batch read, y written
    return {"adjoint": adj, "forward": fwd, "spread": spread, "gather": gather}


def make_inputs(torch, workload, device, seed):
    d, N, m, n, B, C, distribution = WORKLOADS[workload]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if distribution == "uniform":
        pos = torch.rand(n, d, device=device, generator=gen) - 0.5
    else:  # 64 Gaussian clusters, sigma 0.02, wrapped into the torus
        centers = torch.rand(64, d, device=device, generator=gen) * 0.8 - 0.4
        ids = torch.randint(0, 64, (n,), device=device, generator=gen)
        pos = centers[ids] + 0.02 * torch.randn(n, d, device=device, generator=gen)
        pos = ((pos + 0.5) % 1.0) - 0.5
    x = torch.randn(n, C, device=device, generator=gen)
    batch = torch.arange(n, device=device) // (n // B)
    return pos.contiguous(), x.contiguous(), batch.contiguous()


def batch_offsets(torch, n, B, device):
    """batch_size + 1 int64 offsets of the point sets (n / B points each, like make_inputs' batch vector)."""
    return (torch.arange(B + 1, dtype=torch.int64) * (n // B)).to(device)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(kernel_substr):
    """dram read+write bytes per launch of the dominant kernel from the latest committed ncu --set full
    capture (profiles/rNN_roofline.json, written by scripts/summarise_profiles.py); None if absent."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        files = sorted(f for f in os.listdir(pdir) if f.endswith("_roofline.json"))
        with open(os.path.join(pdir, files[-1])) as f:
            data = json.load(f)
        for name, v in data["kernels"].items():
            if kernel_substr in name:
                return v["dram_bytes_per_launch"], "profiles/" + files[-1]
    except Exception:
        pass
    return None, None


def cpu_baseline_ndft(workload, sample_points=192):
    """The reference's only host path: exact NDFT direct sums (reference torch_nfft/ndft.py:5-44),
    timed on this box's host cores on a bounded sample of the workload's point set."""
    import torch
    d, N, m, n, B, C, _ = WORKLOADS[workload]
    kind = "reference"
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "torch_nfft"))
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_ndft", os.path.join(ROOT, "baseline", "_ref", "torch_nfft", "ndft.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        adj, fwd = mod.ndft_adjoint, mod.ndft_forward
    except Exception:
        from oracle import nfft_oracle as O  # port of the same direct sums (numpy)
        kind = "port"
        adj = lambda x, pos, batch, N: torch.from_numpy(O.ndft_adjoint(x.numpy(), pos.numpy(), None, N))
        fwd = lambda y, pos, batch: torch.from_numpy(O.ndft_forward(y.numpy(), pos.numpy(), None))
    g = torch.Generator().manual_seed(0)
    ns = sample_points
    pos = torch.rand(ns, d, generator=g) - 0.5
    x = torch.randn(ns, C, generator=g)
    t0 = time.perf_counter()
    y = adj(x, pos, None, N) if kind == "port" else adj(x, pos, None, N=N)
    f = fwd(y, pos, None)
    dt = time.perf_counter() - t0
    return {"value": ns / dt, "unit": "points/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{ns} of the workload's points (one point set, {C} channel(s)), exact NDFT adjoint+forward "
                      f"on the full {N}^{d} spectrum, {dt:.2f} s of CPU work"}


# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads to the NUMA node of its GPU, so the pinned staging buffers of the
    end-to-end arm are allocated next to the PCIe root the GPU hangs off (matters at 8 ranks per box:
    remote-node staging halves the copy bandwidth).  Returns the node or None when it cannot be found."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id  # older torch: attribute missing
    except Exception:
        bus = None
    try:
        if bus is None:
            import subprocess
            bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
        if isinstance(bus, int):
            return None
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def pick_host_numa_node(torch, dev):
    """When sysfs does not say which NUMA node the GPU hangs off (VMs), measure it: pin this process to each node
    in turn, allocate a pinned buffer there (first touch) and time host->device copies; stay on the fastest node
    for the pinned staging buffers of the end-to-end arm.  Returns (node or None, {node: GB/s})."""
    try:
        nodes = sorted(int(n[4:]) for n in os.listdir("/sys/devices/system/node") if n.startswith("node") and n[4:].isdigit())
    except OSError:
        return None, {}
    allowed0 = os.sched_getaffinity(0)
    cpus = {}
    for node in nodes:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                ids = set()
                for part in f.read().strip().split(","):
                    if part:
                        lo, _, hi = part.partition("-")
                        ids.update(range(int(lo), int(hi or lo) + 1))
            if ids & allowed0:
                cpus[node] = ids & allowed0
        except (OSError, ValueError):
            pass
    if len(cpus) < 2:
        return None, {}
    dst = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
    rates = {}
    for node, ids in cpus.items():
        try:
            os.sched_setaffinity(0, ids)
            src = torch.empty(32 << 20, dtype=torch.uint8).fill_(1).pin_memory()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            rates[node] = 4 * src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
            del src
        except Exception:
            pass
    if not rates:
        os.sched_setaffinity(0, allowed0)
        return None, {}
    best = max(rates, key=rates.get)
    os.sched_setaffinity(0, cpus[best])   # the caller restores the full mask once its pinned buffers exist
    return best, {str(k): round(v, 1) for k, v in rates.items()}


def time_pairs(torch, T, workload, dev, steps, warmup=3, seed=99, graph=False):
    """Device-timed adjoint+forward pairs of another BASELINE workload (inputs resident, one binning per step):
    the `extra_workloads` entries of the default line."""
    d, N, m, n, B, C, distribution = WORKLOADS[workload]
    pos, x, _ = make_inputs(torch, workload, dev, seed=seed)
    ptr = batch_offsets(torch, n, B, dev)

    hint = distribution == "clustered"  # NfftPlan(clustered=True): the device looks for heavy tiles

    def step():
        plan = T.NfftPlan(pos, batch_ptr=ptr, clustered=hint)
        y = T.nfft_adjoint(x, plan=plan, N=N, m=m)
        return T.nfft_forward(y, plan=plan, m=m, real_output=True)

    def run(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms = run(step)
    alg = algorithmic_bytes(d, N, n, B, C)
    peak, _ = measured_peak_gbs()
    out = {"workload": f"{d}D N={N} m={m} n={n} {distribution} B={B} C={C}" + (" (plan hint clustered=True)" if hint else ""),
           "ms_per_step": ms,
           "value": n / (ms * 1e-3), "unit": "points/s", "steps": steps,
           "whole_step_hbm_frac": (alg["adjoint"] + alg["forward"]) / (ms * 1e-3) / 1e9 / peak}
    if graph:
        graphed = T.GraphedTransforms(step)
        ms_g = run(graphed.replay)
        out["cuda_graph"] = {"ms_per_step": ms_g, "value": n / (ms_g * 1e-3), "kernels_per_replay": int(graphed.kernels_per_replay)}
        graphed.close()
    del pos, x, ptr
    torch.cuda.empty_cache()
    return out


def timed_max(torch, dist, dev, fn, steps, warmup, world):
    """K steps bracketed by barrier + synchronize, CUDA events, maximum over the ranks (ms per step)."""
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def strong_c4(torch, dist, T, D, dev, rank, world, steps):
    """The SAME 2^24-point c4 (B = 4 point sets of 2^22) spread over the ranks: whole point sets per rank while
    world <= 4 (no collective); for world = 8 every point set is split over a PAIR of ranks, each spreading its
    half into a partial 256^3 grid, summed by a pairwise all-reduce of the 67 MB grid."""
    d, N, m, n, B, C, _ = WORKLOADS["c4"]
    if world not in (1, 2, 4, 8):
        return None
    per_set = n // B
    split = max(1, world // B)                      # ranks per point set
    sets_here = max(1, B // world)                  # point sets per rank
    group = None
    if split > 1:
        groups = [dist.new_group(list(range(g * split, (g + 1) * split))) for g in range(world // split)]
        group = groups[rank // split]
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    n_here = per_set * sets_here // split
    pos = torch.rand(n_here, d, device=dev, generator=gen) - 0.5
    x = torch.randn(n_here, C, device=dev, generator=gen)
    if split == 1:
        ptr = batch_offsets(torch, n_here, sets_here, dev)

        def step():
            plan = T.NfftPlan(pos, batch_ptr=ptr)
            y = T.nfft_adjoint(x, plan=plan, N=N, m=m)
            return T.nfft_forward(y, plan=plan, m=m, real_output=True)
    else:
        def step():
            plan = T.NfftPlan(pos)  # this rank's slice is binned once per step, for its spread and its gather
            y = D.nfft_adjoint_point_sharded(x, pos, None, N, m, batch_size=1, group=group, plan=plan)
            return D.nfft_forward_point_sharded(y, pos, None, m, True, batch_size=1, group=group, plan=plan)
    ms = timed_max(torch, dist, dev, step, steps, 3, world)
    out = {"workload": "c4: 3D N=128 m=4, 2^24 uniform points in 4 point sets, adjoint+forward", "nranks": world,
           "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "points/s", "scaling": "strong",
           "sharding": f"{sets_here} point set(s) per rank" if split == 1 else
                       f"each point set split over {split} ranks (point sharding)",
           "collective": None}
    if split > 1:
        buf = torch.zeros((2 * N) ** d, device=dev)

        def coll():
            dist.all_reduce(buf, group=group)
        cms = timed_max(torch, dist, dev, coll, 10, 3, world)
        out["collective"] = {"op": f"all_reduce(sum) of the partial real grid within each group of {split} ranks (NCCL)",
                             "bytes": buf.numel() * 4, "ms": cms, "per_step": 1}
    return out


def c5_point_sharded(torch, dist, T, D, dev, rank, world, steps):
    """BASELINE config c5: 3D fastsum with a Gaussian kernel, N=64, m=4, 2^26 points (max-norm 1/4) split over
    the ranks; every rank spreads its slice, ONE all-reduce sums the 128^3 grids, the spectral stage runs on
    every rank, every rank gathers at its own points."""
    n_total, N, m = 2 ** 26, 64, 4
    n = n_total // world
    gen = torch.Generator(device=dev)
    gen.manual_seed(100 + rank)
    pos = (torch.rand(n, 3, device=dev, generator=gen) - 0.5) * 0.5
    x = torch.randn(n, 1, device=dev, generator=gen)
    coeffs = T.gaussian_interpolated_coeffs(0.1, 3, N)

    def step():
        return D.nfft_fastsum_point_sharded(x, coeffs, pos, cutoff=m, batch_size=1)
    ms = timed_max(torch, dist, dev, step, steps, 2, world)
    out = {"workload": "c5: 3D fastsum (Gaussian kernel, sigma 0.1), N=64 m=4, 2^26 points in [-1/4,1/4]^3, symmetric",
           "nranks": world, "ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": "points/s", "scaling": "strong",
           "sharding": "points split evenly over the ranks", "collective": None}
    if world > 1:
        buf = torch.zeros((2 * N) ** 3, device=dev)

        def coll():
            dist.all_reduce(buf)
        cms = timed_max(torch, dist, dev, coll, 10, 3, world)
        out["collective"] = {"op": "all_reduce(sum) of the partial 128^3 real grid over all ranks (NCCL)",
                             "bytes": buf.numel() * 4, "ms": cms, "per_step": 1}
    del pos, x
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch_nfft_b200 as T
    from torch_nfft_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local)   # before any pinned allocation: first touch decides the node
    numa_rates = {}
    cpus_before = os.sched_getaffinity(0)
    if numa is None and not args.no_extras:
        numa, numa_rates = pick_host_numa_node(torch, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, N, m, n, B, C, distribution = WORKLOADS[args.workload]

    pos, x, batch = make_inputs(torch, args.workload, dev, seed=1234 + rank)
    del batch                                     # the point sets are described by B + 1 offsets
    ptr = batch_offsets(torch, n, B, dev)
    # pinned host copies for the end-to-end arm
    h_pos, h_x, h_ptr = (t.cpu().pin_memory() for t in (pos, x, ptr))
    h_spec = torch.empty((B,) + (N,) * d + (C,), dtype=torch.complex64).pin_memory()
    h_y = torch.empty((n, C), dtype=torch.float32).pin_memory()
    if numa_rates:
        os.sched_setaffinity(0, cpus_before)   # the staging buffers are placed: give the CPU baseline all cores back

    def pair(x_, pos_, ptr_):
        # every step is a fresh transform pair: its points are binned once (NfftPlan) and the adjoint and the
        # forward transform of the pair share that binning, as forward + backward of autograd do
        plan = T.NfftPlan(pos_, batch_ptr=ptr_, clustered=distribution == "clustered")
        y = T.nfft_adjoint(x_, plan=plan, N=N, m=m)
        return y, T.nfft_forward(y, plan=plan, m=m, real_output=True)

    def step_device():
        return pair(x, pos, ptr)[1]

    # End-to-end arm: every step copies ITS inputs from pinned host memory and reads ITS results back.
    # Like a data loader, the copies of step k+1 (copy stream) overlap the transforms of step k
    # (compute stream) through two device buffer sets; results leave on a third stream.
    compute_stream = torch.cuda.current_stream(dev)
    h2d_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dbuf = [(torch.empty_like(pos), torch.empty_like(x), torch.empty_like(ptr)) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]   # inputs of the buffer set have arrived
    ev_free = [torch.cuda.Event() for _ in range(2)]    # the transforms reading the buffer set are done
    e2e_state = {"k": 0, "staged": False, "keep": None}

    def stage_inputs(slot):
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(ev_free[slot])
            dpos, dx, dptr = dbuf[slot]
            dpos.copy_(h_pos, non_blocking=True)
            dx.copy_(h_x, non_blocking=True)
            dptr.copy_(h_ptr, non_blocking=True)
            ev_ready[slot].record(h2d_stream)

    def step_e2e():
        k = e2e_state["k"]
        slot = k % 2
        if not e2e_state["staged"]:
            stage_inputs(slot)                 # first step: nothing to overlap with
        stage_inputs(1 - slot)                 # inputs of the NEXT step travel while this one computes
        e2e_state["staged"] = True
        compute_stream.wait_event(ev_ready[slot])
        dpos, dx, dptr = dbuf[slot]
        y, f = pair(dx, dpos, dptr)
        ev_free[slot].record(compute_stream)
        done = torch.cuda.Event()
        done.record(compute_stream)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            h_spec.copy_(y, non_blocking=True)
            h_y.copy_(f, non_blocking=True)
        y.record_stream(d2h_stream)
        f.record_stream(d2h_stream)
        # the step is complete only when its results are on the host: the compute stream (which the
        # timing events are recorded on) waits for this step's read-back before the next step ends
        back = torch.cuda.Event()
        back.record(d2h_stream)
        e2e_state["keep"] = (y, f, back)
        e2e_state["k"] = k + 1

    def finish_e2e():
        if e2e_state["keep"] is not None:
            compute_stream.wait_event(e2e_state["keep"][2])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    def timed_e2e(step, finish, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        finish()           # the compute stream waits for the last read-back
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), 0, 0

    sampler = ClockSampler(local) if rank == 0 else None  # started early: nvidia-smi needs ~0.3 s to report
    for _ in range(max(args.warmup, 3)):
        step_device()
    _lib.profile_enable(True)
    _lib.profile_read()
    launches0 = _lib.launch_count()
    ms_total, t0, t1 = timed(step_device, args.steps)
    launches = _lib.launch_count() - launches0
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    clocks = sampler.stop(t0, t1) if sampler else None

    graph_info = None
    if args.cuda_graph:
        # the same step captured once into a CUDA graph and replayed (no per-launch host cost: matters for
        # the launch-bound 1D workload c2, 14 kernels per pair); reported next to the eager numbers
        graphed = T.GraphedTransforms(step_device)
        graph, captured = graphed.graph, graphed.kernels_per_replay
        for _ in range(3):
            graph.replay()
        ms_graph, _, _ = timed(graph.replay, args.steps)
        graph_info = {"ms_per_step": ms_graph / args.steps, "value": n * world / (ms_graph / args.steps * 1e-3),
                      "unit": "points/s", "kernels_per_replay": int(captured)}
        del graph, graphed

    if args.no_extras:
        ms_e2e = float("nan")
    else:
        for e in ev_free:
            e.record(compute_stream)
        for _ in range(2):
            step_e2e()
        finish_e2e()
        torch.cuda.synchronize()

        # K timed steps; K+1 input sets are copied (the last prefetch is extra work inside the region)
        e2e_state["k"] = 2
        ms_e2e, _, _ = timed_e2e(step_e2e, finish_e2e, args.steps)

    # the other BASELINE configurations (device-timed, inputs resident), rank 0 at N = 1 only
    extra = None
    multi = {}
    if not args.no_extras:
        del dbuf
        torch.cuda.empty_cache()
        if world == 1 and args.workload == "c4":
            ks = max(3, min(args.steps, 10))
            extra = {"c4_clustered": time_pairs(torch, T, "c4_clustered", dev, ks),
                     "c3": time_pairs(torch, T, "c3", dev, ks),
                     "c2": time_pairs(torch, T, "c2", dev, 50, warmup=5, graph=True)}
        if args.workload == "c4":
            from torch_nfft_b200 import dist as D
            ks = max(3, min(args.steps, 10))
            if world > 1:
                del pos, x
                torch.cuda.empty_cache()
                multi["strong_c4"] = strong_c4(torch, dist, T, D, dev, rank, world, ks)
            multi["c5_point_sharded"] = c5_point_sharded(torch, dist, T, D, dev, rank, world, max(3, min(args.steps, 5)))

    if rank == 0:
        ms_step = ms_total / args.steps
        alg = algorithmic_bytes(d, N, n, B, C)
        peak, peak_src = measured_peak_gbs()
        spread_ms = prof["spread"][0] / max(prof["spread"][1], 1)
        gather_ms = prof["gather"][0] / max(prof["gather"][1], 1)
        achieved = alg["spread"] / (spread_ms * 1e-3) / 1e9
        stage_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]}
        taps = n * C * (2 * m + 2) ** d
        traffic, traffic_src = profiled_traffic("spread")
        out = {
            "metric": "NU points/sec (adjoint+forward)",
            "value": n * world / (ms_step * 1e-3),
            "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {d}D adjoint+forward NFFT, N={N}, m={m}, n={n} {distribution} "
                                   f"points per GPU, batch_size={B}, {C} channel(s), real x -> complex spectrum -> real y",
                       "sharding": "batch entries per GPU, no collective" if world > 1 else "single GPU",
                       "binning": "once per step (NfftPlan made inside the step, shared by its adjoint and forward); "
                                  "nothing is reused across steps",
                       "point_sets": "batch_ptr: batch_size + 1 int64 offsets (the per-point int64 batch vector of the "
                                     "reference is accepted too)",
                       "l2": "inputs larger than L2 (pos+x = %d MB per step, grid %d MB)" % (
                           (n * (4 * d + 4 * C)) >> 20, (B * C * (2 * N) ** d * 4) >> 20)},
            "e2e": {"value": n * world / (ms_e2e / args.steps * 1e-3), "unit": "points/s",
                    "h2d_bytes_per_step": n * (4 * d + 4 * C) + 8 * (B + 1),
                    "d2h_bytes_per_step": B * C * N ** d * 8 + n * C * 4,
                    "host_numa_node_rank0": numa, "h2d_gbs_by_numa_node_rank0": numa_rates},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "spread kernel (adjoint window convolution; spread_reg_kernel at c4)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg["spread"], "ms_per_launch": spread_ms,
                         "whole_step": {"algorithmic_bytes": alg["adjoint"] + alg["forward"],
                                        "achieved": (alg["adjoint"] + alg["forward"]) / (ms_step * 1e-3) / 1e9,
                                        "frac": (alg["adjoint"] + alg["forward"]) / (ms_step * 1e-3) / 1e9 / peak},
                         "gather": {"achieved": alg["gather"] / (gather_ms * 1e-3) / 1e9 if gather_ms else None,
                                    "ms_per_launch": gather_ms},
                         "taps_per_s": {"spread": taps / (spread_ms * 1e-3), "gather": taps / (gather_ms * 1e-3) if gather_ms else None}},
            "stage_ms_per_step": stage_ms,
            "cpu_baseline": cpu_baseline_ndft(args.workload) if (world == 1 and not args.no_extras) else None,
        }
        if graph_info is not None:
            out["cuda_graph"] = graph_info
        if extra is not None:
            out["extra_workloads"] = extra
        out.update(multi)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The unmodified reference (baseline/_ref): its CUDA NFFT through its own public API on the
    same config (north_star: "next to the reference's CUDA NFFT on the same B200"), together with
    its only CPU path (ndft) as cpu_baseline.  Falls back to the CPU path alone without CUDA."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    d, N, m, n, B, C, distribution = WORKLOADS[args.workload]
    cpu = cpu_baseline_ndft(args.workload)
    base = {"impl": "reference", "metric": "NU points/sec (adjoint+forward)", "unit": "points/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {d}D adjoint+forward NFFT, N={N}, m={m}, n={n} {distribution} points, "
                                   f"batch_size={B}, {C} channel(s)"},
            "cpu_baseline": cpu}
    ref = None
    if args.ref_device == "cuda" and torch.cuda.is_available():
        try:
            sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
            import torch_nfft as ref  # noqa
        except Exception as e:  # pragma: no cover
            base["reference_cuda_unavailable"] = repr(e)[:200]
            ref = None
    if ref is None:
        base.update({"value": cpu["value"], "ms_per_step": None, "gpu_launches": 0,
                     "e2e": {"value": cpu["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "config": dict(base["config"], arm="reference CPU path (ndft direct sums), bounded sample")})
        print(json.dumps(base))
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pos, x, batch = make_inputs(torch, args.workload, dev, seed=1234)
    # The reference needs ~23 s per adjoint+forward pair at the full 2^24 points (one global atomic per tap).
    # The arm runs the FULL workload (same_config) whenever steps + warm-up fit a ~12 minute run -- the driver's
    # `--steps 20 --warmup 5` does (25 x 24 s = 10 min).  Only longer requests fall back to a bounded sample
    # (every `stride`-th point of each point set: same grid, n / stride points; its cost is linear in n).
    budget_s = 720.0
    if args.ref_points > 0:
        ref_points = args.ref_points
    else:
        # measured on B200: about 0.85 s fixed + 1.38 us per point per adjoint+forward pair
        ref_points = n
        while ref_points > 2 ** 18 and (0.85 + 1.38e-6 * ref_points) * (args.steps + max(args.warmup, 1)) > budget_s:
            ref_points //= 2
    stride = max(1, n // ref_points)
    if stride > 1:
        pos, x, batch = pos[::stride].contiguous(), x[::stride].contiguous(), batch[::stride].contiguous()
    ns = pos.shape[0]

    def step():
        y = ref.nfft_adjoint(x, pos, batch, N, m)
        return ref.nfft_forward(y, pos, batch, m, True)

    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    v = ns / (ms_step * 1e-3)
    base.update({"value": v, "ms_per_step": ms_step, "gpu_launches": 0,
                 "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "config": dict(base["config"], arm="reference CUDA NFFT (baseline/_ref, torch_nfft.nfft_adjoint + "
                                                    "nfft_forward), inputs resident on the GPU",
                                sample=("the full workload" if stride == 1 else
                                        f"{ns} of the {n} points per step (every {stride}-th point of each point set)")
                                       + f", full grid N={N}, batch_size={B}", same_config=stride == 1)})
    print(json.dumps(base))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--cuda-graph", action="store_true",
                    help="also time the step captured into a CUDA graph (adds a `cuda_graph` object to the line)")
    ap.add_argument("--no-extras", action="store_true", help="development: skip the e2e and cpu_baseline legs")
    ap.add_argument("--ref-points", type=int, default=0,
                    help="points per step of the reference CUDA arm (bounded sample of the workload); 0 = as many "
                         "as fit a ~4 minute run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
