"""Turn an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) of ONE kernel launch into the short metric list kept under
profiles/.  usage: python tools/ncu_summary.py RAW.csv "header line" > profiles/rNN_x_ncu_summary.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[-1]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gcc__average_cache_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "local_load", "local_store", "Kernel Name")
for line in sys.argv[2:]:
    print(line)
print()
for h, u, v in sorted(zip(hdr, units, vals)):
    if h in keep or h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") or "local" in h and h.startswith("smsp__inst_executed_op"):
        print(f"{h:<95s} {v} {u}")
