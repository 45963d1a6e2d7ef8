"""Machine-readable evidence from `ncu --set full` reports: selected raw metrics per kernel as CSV (+ JSON for bench.py)
and the hottest source lines by warp-stall samples.

    python tools/ncu_extract.py profiles/r02 gpurun_out/r02x_ncu_reg_layer_fused.ncu-rep [more.ncu-rep ...]

writes  profiles/r02_ncu_metrics.csv, profiles/r02_ncu_metrics.json and profiles/r02_ncu_source_<kernel>.csv (top 40 lines).
bench.py reads `roofline.traffic` from the JSON (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel)."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__pipe_tensor.*|sm__inst_executed_pipe_tensor.*|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__inst_executed\.sum|sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|"
    r"launch__block_size|launch__grid_size|launch__shared_mem_per_block_dynamic|launch__occupancy_limit.*|"
    r"sm__cycles_elapsed\.max|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|l1tex__data_bank_conflicts.*|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|derived__smsp__inst_executed_op_.*|"
    r"smsp__inst_executed_op_local.*|sm__sass_inst_executed_op_local.*)$")


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_page(rep):
    """Per-source-line aggregates of the `cuda,sass` view: [(file, line, text, samples, instructions executed)]."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    lines, fname, col_s, col_i = [], "?", None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            col_s, col_i = r.index("# Samples"), r.index("Instructions Executed")
        elif col_s is not None and r[0].isdigit():
            try:
                lines.append((fname, int(r[0]), r[1].strip(), float(r[col_s]), float(r[col_i])))
            except ValueError:
                pass
    return lines


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    table, js = [], {}
    for rep in reps:
        hdr, units, launches = raw_page(rep)
        for vals in launches:
            d = dict(zip(hdr, vals))
            name = re.sub(r"\(.*", "", d.get("Kernel Name", "?")).replace("void ", "")
            short = re.sub(r"<.*", "", name).split("::")[-1]
            js.setdefault(short, {"kernel": name, "report": rep.split("/")[-1]})
            for h, u, v in zip(hdr, units, vals):
                if KEEP.match(h):
                    table.append((name, h, u, v))
                    try:
                        js[short][h] = float(v.replace(",", ""))
                        js[short][h + "#unit"] = u
                    except ValueError:
                        pass
        src = source_page(rep)
        if src:
            tot = sum(x[3] for x in src) or 1.0
            short = re.sub(r"<.*", "", re.sub(r"\(.*", "", launches[0][hdr.index("Kernel Name")]).replace("void ", "")).split("::")[-1]
            with open(f"{prefix}_ncu_source_{short}.csv", "w", newline="") as f:
                wr = csv.writer(f)
                wr.writerow(["share_of_stall_samples", "samples", "warp_instructions", "file", "line", "source"])
                for x in sorted(src, key=lambda x: -x[3])[:60]:
                    wr.writerow(["%.4f" % (x[3] / tot), int(x[3]), int(x[4]), x[0], x[1], x[2][:160]])
    with open(f"{prefix}_ncu_metrics.csv", "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["kernel", "metric", "unit", "value"])
        wr.writerows(table)
    json.dump(js, open(f"{prefix}_ncu_metrics.json", "w"), indent=1)
    for k, d in js.items():
        mb = lambda key: d.get(key, 0.0) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(d.get(key + "#unit", "byte"), 1.0)
        tens = [v for kk, v in d.items() if kk.startswith("sm__pipe_tensor") and kk.endswith("pct_of_peak_sustained_active") and isinstance(v, float)]
        print(f"{k}: {d.get('gpu__time_duration.sum')} {d.get('gpu__time_duration.sum#unit')}, dram {mb('dram__bytes_read.sum') / 1e6:.1f} + "
              f"{mb('dram__bytes_write.sum') / 1e6:.1f} MB, tensor pipe active {max(tens) if tens else None} %, issue active "
              f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} %, regs {d.get('launch__registers_per_thread')}")


if __name__ == "__main__":
    main()
