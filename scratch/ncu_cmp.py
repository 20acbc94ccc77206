import csv, sys
want = ["gpu__time_duration.sum","sm__cycles_elapsed.max","gpc__cycles_elapsed.avg.per_second",
"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
"sm__inst_executed_pipe_tc.sum","sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
"l1tex__data_pipe_tc_wavefronts_mem_shared.sum","l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
"lts__throughput.avg.pct_of_peak_sustained_elapsed","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_bytes.sum","dram__bytes_read.sum","dram__bytes_write.sum",
"l1tex__m_xbar2l1tex_read_bytes.sum","l1tex__m_xbar2l1tex_read_bytes.sum.per_second","sm__cycles_active.avg","smsp__inst_executed.sum",
"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__warps_active.avg.pct_of_peak_sustained_active","dram__throughput.avg.pct_of_peak_sustained_elapsed",
"lts__t_sector_hit_rate.pct","sm__sass_inst_executed_op_shared_st.sum"]
tabs = {}
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    tabs[f] = dict(zip(rows[0], zip(rows[1], rows[2])))
allk = list(next(iter(tabs.values())).keys())
extra = [k for k in allk if ("tensor" in k or "pipe_tc" in k or "tmem" in k.lower() or "utc" in k.lower()) and k not in want and "pct" in k and ".avg." in k]
for k in want + extra:
    vals = [tabs[f].get(k, ("", "n/a")) for f in sys.argv[1:]]
    print(f"{k[:88]:88s} {vals[0][0]:10s} " + " ".join(f"{v[1][:16]:>16s}" for v in vals))
