"""Key metrics per launch from an `ncu -i X.ncu-rep --page raw --csv` export.
    python profiles/summarize_ncu.py gpurun_out/r02_prof_pt_v2.csv [more.csv ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tensor_subpipe_umma",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sector_op_read_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_tma.sum", "sm__inst_executed_pipe_tma.sum"]

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        print(f"{d.get('Kernel Name', '?')[:110]} grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for i, n in enumerate(names):
            utc = "utc" in n.lower() and r[i] not in ("0", "", "n/a") and "peak_sustained" not in n.replace("pct_of_peak_sustained", "") \
                and (n.endswith(".sum") or n.endswith("avg.pct_of_peak_sustained_elapsed"))
            if n in KEYS or utc:
                print(f"   {n:110s} {r[i]:>16s} {units[i]}")
        print()
