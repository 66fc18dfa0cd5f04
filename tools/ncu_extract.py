"""Extract the headline metrics of an ncu --set full report (.ncu-rep) into a small text summary for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for h in sys.argv[2:]:
    print("# " + h)
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"kernel = {d.get('Kernel Name')}  grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}")
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"{k} = {d[k]} {units[hdr.index(k)]}")
