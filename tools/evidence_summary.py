"""profiles/r03_summary.md from the committed evidence files (bench line, launch lists, ncu exports).
    python tools/evidence_summary.py > profiles/r03_summary.md"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda f: os.path.join(ROOT, "profiles", f)
d = json.loads(open(P("r03_bench_final.json")).read().strip().splitlines()[-1])
m = json.load(open(P("r03_ragged_ncu_metrics.json")))
md = json.load(open(P("r03_dense_ncu_metrics.json")))
out = []
out.append("# Round 2, final build: evidence summary (B200, sm_100a)\n")
out.append("Files: `r03_bench_final.json` (bench line of `python bench.py`), `r03_bench_reference_arm.json` (`--impl reference`), `r03_bench_4gpu.json` (torchrun, 4 ranks),")
out.append("`r03_launches_bf16_infer18955_{dense,ragged}.csv` (`ncu --metrics gpu__time_duration.sum --clock-control none`, second pass = warm), `r03_dense_ncu_metrics.{csv,json}` /")
out.append("`r03_ragged_ncu_metrics.{csv,json}` and `r03_*_ncu_source_*.csv` (`ncu --set full --import-source on` of one launch each), `r03_sass.txt` (mnemonic counts of the built library).\n")
out.append("## Bench line (1 GPU)\n")
out.append("| | |\n|---|---|")
out.append(f"| `value` (18,955 dense genes resident in HBM, one launch chain of {d['gpu_launches_per_step']} kernels) | {d['value']/1e6:.2f} M genes/s, {d['ms_per_step']:.2f} ms per sweep |")
e = d["e2e"]
out.append(f"| `e2e` (pinned host -> logits on the host, zero-suppressed FP16 wire) | {e['value']/1e6:.2f} M genes/s, {e['h2d_bytes_per_gene']/1e3:.1f} kB per gene, {e['h2d_gbs']:.1f} GB/s of host->device copies |")
for k, v in e["variants"].items():
    out.append(f"| e2e variant `{k}` | {v['value']/1e6:.3f} M genes/s |")
r = d["ragged"]
out.append(f"| ragged sweep with the plan | {r['value']/1e6:.2f} M genes/s ({r['ms_per_step']:.2f} ms), {r['vs_dense']:.2f} x the dense rate; without the plan {r['without_plan']['value']/1e6:.2f} M; dense batch forced through the plan {r['dense_batch_with_plan']['value']/1e6:.2f} M |")
ro = d["roofline"]
out.append(f"| dominant kernel (`reg_layer_fused_kernel<9>`, timed alone over the 18,955-gene launch) | {ro['ms_per_launch']:.3f} ms = {ro['achieved']:.0f} TFLOP/s = {100*ro['frac']:.1f} % of the measured BF16 peak ({ro['peak']:.0f} TFLOP/s); DRAM traffic {ro['traffic']['bytes']/1e6:.0f} MB against {ro['algorithmic_bytes_per_launch']/1e6:.0f} MB algorithmic |")
sq = d["sqa_kernel"]
out.append(f"| `sqa_fused_kernel`, n = 400 alone | {sq['achieved']:.0f} GB/s = {100*sq['frac']:.1f} % of the measured HBM peak |")
rp = d["raw_depth_path"]
out.append(f"| raw FP16 depth in HBM -> logits | {rp['value']/1e6:.2f} M genes/s = {rp['hbm_gbs']:.0f} GB/s of raw depth (one stream: {rp['one_stream']['value']/1e6:.2f} M) |")
out.append(f"| 44 checkpoints x 18,955 genes | {d['ensemble_sweep']['ms']:.0f} ms |")
t = d["train"]
out.append(f"| training step (Chromoformer-reg, batch 64, BF16 contractions) | {t['ms_per_step']:.3f} ms = {t['value']/1e3:.1f} k samples/s, {t['gpu_launches']} launches; reference on the host cores {t['cpu_baseline']['value']:.0f} samples/s |")
out.append(f"| CPU baseline (unmodified reference, {d['cpu_baseline']['cores']} host cores) | {d['cpu_baseline']['value']:.0f} genes/s |")
out.append(f"| clocks during the timed region | {d['clocks']['sm_mhz']} / {d['clocks']['sm_max_mhz']} MHz, reasons {d['clocks']['reasons']} |\n")


def chain(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    idx = max(i for i, r in enumerate(rows) if "centre_embed" in r[4])          # the second (warm) forward
    agg = {}
    for r in rows[idx:]:
        name = r[4].split("(")[0].replace("chromo::", "").replace("<unnamed>::", "").replace("void ", "")
        if "cub::" in name:
            name = "cub radix sort (3 kernels)"
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += float(r[-1]) / 1e3
    tot = sum(v[1] for v in agg.values())
    return tot, sorted(agg.items(), key=lambda kv: -kv[1][1])


out.append("## Launch chain of one sweep (second, warm pass; times are ncu's serialised per-launch durations)\n")
for name in ("dense", "ragged"):
    tot, items = chain(P(f"r03_launches_bf16_infer18955_{name}.csv"))
    out.append(f"**{name}** ({'CHROMO_F_DENSE hint, no plan' if name == 'dense' else 'ragged plan'}): {sum(v[0] for _, v in items)} launches, {tot/1e3:.3f} ms\n")
    out.append("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, v in items:
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100*v[1]/tot:.1f} % |")
    out.append("")
out.append("## `ncu --set full` (one launch each)\n")
out.append("| kernel | sweep | duration | DRAM read + write | tensor pipe active | issue active | registers |\n|---|---|---:|---:|---:|---:|---:|")
for label, mm in (("dense", md), ("ragged", m)):
    for k in mm:
        x = mm[k]
        out.append(f"| `{k}` | {label} | {x['gpu__time_duration.sum']:.3f} {x['gpu__time_duration.sum#unit']} | {x['dram__bytes_read.sum']:.0f} + {x['dram__bytes_write.sum']:.0f} MB | "
                   f"{x['sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active']:.1f} % | {x['smsp__issue_active.avg.pct_of_peak_sustained_active']:.1f} % | {x.get('launch__registers_per_thread', '')} |")
out.append("\nStall breakdowns and the hottest source lines per kernel: `r03_*_ncu_metrics.csv`, `r03_*_ncu_source_*.csv`.")
print("\n".join(out))
