"""ncu -i <rep> --page raw --csv | python tools/ncu_extract.py <label>  ->  one compact block of the metrics the roofline needs."""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__sass_inst_executed_op_utcmma.sum", "smsp__inst_executed_op_tma_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active"]
rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r]
if not hi:
    print("#", sys.argv[1], "NO DATA")
    sys.exit(0)
H, U = rows[hi[0]], rows[hi[0] + 1]
for V in rows[hi[0] + 2:]:
    if len(V) != len(H):
        continue
    print("---", sys.argv[1])
    print("Kernel Name =", V[H.index("Kernel Name")][:160])
    for k in KEEP:
        if k in H:
            i = H.index(k)
            print(f"{k} [{U[i]}] = {V[i]}")
