#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep OUT.csv [--traffic profiles/traffic.json]
Selected metrics of an `ncu --set full` capture (read with `ncu -i ... --page raw --csv`), one column per captured
launch -- the table committed under profiles/ -- and, optionally, the DRAM bytes per launch bench.py reports as
roofline.traffic."""
import csv
import io
import json
import subprocess
import sys

KEEP_PREFIX = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
               "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
               "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
               "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
               "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
               "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
               "dram__bytes_read.sum.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
               "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
               "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
               "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
               "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
               "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
               "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
               "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units, data = rows[0], rows[1], rows[2:]
    names = [r[header.index("Kernel Name")] for r in data]
    keep = [i for i, h in enumerate(header) if h in KEEP_PREFIX or
            (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch {k + 1} ({n.split('(')[0][:40]})" for k, n in enumerate(names)])
        for i in keep:
            w.writerow([header[i], units[i]] + [r[i] for r in data])
    print("wrote", out, "for", len(data), "launches")
    if "--traffic" in sys.argv:
        tp = sys.argv[sys.argv.index("--traffic") + 1]

        def col(name):
            i = header.index(name)
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
            return [float(r[i].replace(",", "")) * scale for r in data]
        per = [a + b for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
        json.dump({"env_step64_kernel_bytes_per_launch": sum(per) / len(per),
                   "source": f"{out} (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, mean of {len(per)} "
                             "launches after 100 warm-up steps)"}, open(tp, "w"), indent=1)
        print("wrote", tp, sum(per) / len(per))


if __name__ == "__main__":
    main()
