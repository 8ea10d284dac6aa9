"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name')}  (id {d.get('ID')})")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:78s} {d[k]:>18s} {u.get(k, '')}")
        stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k]), reverse=True)
        print("  warp stall reasons (avg warps stalled per issue-active cycle):")
        for v, k in stalls[:8]:
            print(f"    {k[len(STALL):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
        try:
            fl = sum(float(d[k].replace(",", "")) * m for k, m in (("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 1),
                     ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 1), ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 2)))
            t = float(d["gpu__time_duration.sum"].replace(",", ""))
            tu = u["gpu__time_duration.sum"]
            sec = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
            print(f"  executed FP64 flop (dadd + dmul + 2 dfma) = {fl:.4g}  -> {fl / sec / 1e12:.2f} TFLOP/s under the profiler")
        except Exception as e:  # noqa: BLE001
            print("  (fp64 flop count unavailable:", e, ")")


if __name__ == "__main__":
    main(sys.argv[1])
