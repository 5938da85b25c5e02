"""Key metrics of one `ncu --set full` report: python scripts/ncu_summary.py report.ncu-rep
The tensor-pipe counter that tracks tcgen05.mma (SASS UTCHMMA) on sm_100 is
sm__ops_path_tensor_op_utchmma_src_<type>_dst_fp32_sparsity_off (the legacy sm__pipe_tensor_cycles_active_realtime
counter, which the report carries with a TPC.TriageCompute. prefix, only sees a few per cent of it)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tma.sum", "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_utcmma.sum",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_read.sum"]
suffix = ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",)      # prefixed in the report
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        tensor = h.startswith("sm__ops_path_tensor_op_utchmma") and (h.endswith(".sum") or h.endswith(".avg.pct_of_peak_sustained_elapsed")) and r[i] not in ("0", "")
        if h in want or h.startswith("launch__cluster") or tensor or h.endswith(suffix):
            print(f"{h:100s} {units[i]:12s} {r[i]}")
