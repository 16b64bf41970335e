"""Builds library variants for an A/B session into gpurun_variants/ (git-ignored, travels with gpurun).

    python scripts/build_variants.py                 # the queued round-2 experiments below
    python scripts/build_variants.py name:D1=1,D2=4  # explicit variants

Then on the GPU:  gpurun -- 'bash scripts/final_round.sh r02a c4'   (A/B, then tests / bench / ncu with
the winner) or 'bash scripts/quick_ab.sh r02a c4'.  A variant named *base* is the reference point.
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_nfft_b200 import _build  # noqa: E402

QUEUED = {
    "a_base": [],
    "b_walk8": ["NFFT_REG_WALK_UNROLL=8"],          # flush: 8 reductions in flight per thread (4 gave -3.6 %)
    "c_ttas": ["NFFT_REG_LOCK_TTAS=1"],             # add-out lock: poll with a load, CAS only when free
    "d_ttas_ns8": ["NFFT_REG_LOCK_TTAS=1", "NFFT_REG_LOCK_NS=8"],
    "e_lockns8": ["NFFT_REG_LOCK_NS=8"],
    "f_maxpts1280": ["NFFT_REG_MAXPTS=1280"],       # smaller point buffer: more shared memory head-room
}


def main():
    variants = dict(QUEUED)
    if len(sys.argv) > 1:
        variants = {}
        for arg in sys.argv[1:]:
            name, _, defs = arg.partition(":")
            variants[name] = [d for d in defs.split(",") if d]
    out = os.path.join(ROOT, "gpurun_variants")
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(out):
        if f.startswith("lib_") and f.endswith(".so"):
            os.remove(os.path.join(out, f))

    def one(item):
        name, defs = item
        path = os.path.join(out, f"lib_{name}.so")
        _build.build(force=True, defines=defs, out=path)
        return name, defs

    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, defs in ex.map(one, variants.items()):
            print(f"built lib_{name}.so  {' '.join('-D' + d for d in defs)}")


if __name__ == "__main__":
    main()
