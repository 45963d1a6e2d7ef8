"""Kernel-share table from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/summarize_launches.py profiles/r01_launches_bf16_v10_infer4096.csv [marker] [pass_index]

Rows before the `pass_index`-th occurrence of `marker` (default: the 2nd `centre_embed`, i.e. the warm pass
of tools/profile_forward.py) are dropped."""
import collections
import csv
import re
import sys


def load(path, marker="centre_embed", which=1):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"]) for r in rows]
    idx = [i for i, (n, _, _) in enumerate(names) if marker in n]
    return names[idx[which]:] if len(idx) > which else names


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "centre_embed"
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sec = load(path, marker, which)
    tot = sum(t for _, t, _ in sec)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in sec:
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("chromo::", "")
        agg[k][0] += 1
        agg[k][1] += t
    print(f"{len(sec)} launches, {tot / 1e6:.3f} ms serialised\n")
    print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
