import csv, sys
keys = ["gpu__time_duration.sum","sm__cycles_elapsed.max","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tc","lts__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","lts__t_sectors_srcunit_tex_op_read.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared","smsp__cycles_active.avg","sm__throughput.avg.pct","l1tex__throughput.avg.pct","sm__cycles_active.avg","launch__grid_size","launch__cluster","smsp__pcsamp_warps_issue_stalled","sm__clock","gpc__cycles_elapsed.avg.per_second","smsp__warp_issue_stalled","sm__mem","shared"]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("==", f)
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) or k in h for k in keys):
            print(f"  {h:95s} {u:12s} {v}")
