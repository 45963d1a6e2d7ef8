"""Stage the UNMODIFIED reference demo (script + the data of the first N demo genes) under baseline/_ref/demo
so that tests/test_reference_demo_gpu.py can run `demo/run_demo.py` itself against this package on the GPU box.
baseline/_ref is git-ignored (nothing of the reference enters the history) but travels with gpurun snapshots.

    python tools/stage_reference_demo.py [N=24]
"""
import os
import shutil
import sys

import pandas as pd

REF = os.environ.get("CHROMOFORMER_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref", "demo")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    os.makedirs(os.path.join(DST, "demo_data"), exist_ok=True)
    for f in ("run_demo.py", "run_demo_regression.py", "random_prediction.out"):
        shutil.copy(os.path.join(REF, "demo", f), os.path.join(DST, f))
    meta = pd.read_csv(os.path.join(REF, "demo", "demo_meta.csv")).iloc[:n]
    meta.to_csv(os.path.join(DST, "demo_meta_head.csv"), index=False)
    files = set()
    for r in meta.itertuples(index=False):
        files.add(f"{r.chrom}:{r.start - 20000}-{r.start + 20000}.npy")
        if isinstance(r.neighbors, str):
            files.update(f"{x}.npy" for x in r.neighbors.split(";") if x)
    size = 0
    for f in sorted(files):
        src = os.path.join(REF, "demo", "demo_data", f)
        shutil.copy(src, os.path.join(DST, "demo_data", f))
        size += os.path.getsize(src)
    print(f"staged {len(meta)} genes, {len(files)} region files, {size / 1e6:.1f} MB -> {DST}")


if __name__ == "__main__":
    main()
