"""One line of key metrics per kernel of an .ncu-rep (run where ncu is installed):
python tests/micro/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r1/x_summary.txt"""
import csv, subprocess, sys

KEYS = [("launch__grid_size", "launch__grid_size"), ("launch__block_size", "launch__block_size"),
        ("gpu__time_duration.sum", "gpu__time_duration"), ("dram__bytes_read.sum", "dram__bytes_read"),
        ("dram__bytes_write.sum", "dram__bytes_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput"),
        ("lts__t_sector_hit_rate.pct", "lts__t_sector_hit_rate"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_hmma_cycles_active"),
        ("launch__registers_per_thread", "launch__registers_per_thread"),
        ("launch__occupancy_limit_registers", "launch__occupancy_limit_registers"),
        ("launch__occupancy_limit_shared_mem", "launch__occupancy_limit_shared_mem"),
        ("smsp__inst_executed.sum", "smsp__inst_executed"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared"),
        ("sm__cycles_elapsed.max", "sm__cycles_elapsed")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
ci = {x: i for i, x in enumerate(h)}
for r in rows[2:]:
    print(r[ci["Kernel Name"]][:110])
    print("  " + " | ".join("%s=%s" % (short, r[ci[k]]) for k, short in KEYS if k in ci))
