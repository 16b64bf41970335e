"""Runs the UNMODIFIED compiled reference (baseline/_ref, dominikbuenger/torch_nfft) on one seeded case.

Executed as a subprocess by tests/test_parity_reference_gpu.py: the reference exit()s the process on
CUDA errors (csrc/cuda/cuda_utils.cu:7-14), its package and op namespace are both called `torch_nfft`,
and its cuFFT runs on the legacy stream -- none of which should live in the test process.

    python tests/ref_runner.py <case> <out.pt>

Inputs come from tests/fullsize_cases.py (CPU generator, seeded: bit-identical in both processes).
Writes {"y": adjoint spectrum, "f": forward values at the same points computed FROM that spectrum,
"run_to_run": rel-L2 between two reference adjoint runs (its float-atomic noise floor), "seconds": ...}
or, for fastsum cases, {"s": fastsum result, ...}.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def main():
    case, out = sys.argv[1], sys.argv[2]
    import torch_nfft as ref  # the reference
    from fullsize_cases import make_case

    c = make_case(case)
    dev = torch.device("cuda")
    pos, x, batch = c["pos"].to(dev), c["x"].to(dev), c["batch"].to(dev)
    N, m = c["N"], c["m"]
    res = {}
    t0 = time.time()
    if c["op"] == "pair":
        y = ref.nfft_adjoint(x, pos, batch, N, m)                      # reference nfft.py:31
        y2 = ref.nfft_adjoint(x, pos, batch, N, m)
        res["run_to_run"] = float((torch.linalg.vector_norm(y - y2) / torch.linalg.vector_norm(y)).item())
        del y2
        f = ref.nfft_forward(y, pos, batch, m, True)                    # reference nfft.py:57
        res["y"], res["f"] = y.cpu(), f.cpu()
    else:
        coeffs = c["coeffs"].to(dev)
        s = ref.nfft_fastsum(x, coeffs, pos, batch=batch, cutoff=m)    # symmetric path (core_cuda.cu:552)
        s2 = ref.nfft_fastsum(x, coeffs, pos, batch=batch, cutoff=m)
        res["run_to_run"] = float((torch.linalg.vector_norm(s - s2) / torch.linalg.vector_norm(s)).item())
        res["s"] = s.cpu()
    torch.cuda.synchronize()
    res["seconds"] = time.time() - t0
    torch.save(res, out)


if __name__ == "__main__":
    main()
