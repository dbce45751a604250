"""Turn an .ncu-rep into a small text summary for profiles/: per kernel launch the headline metrics
(duration, registers, occupancy, pipe utilisation, DRAM bytes, stall reasons) and the SASS-region
breakdown of tools/ncu_regions.py.   usage: ncu_summary.py report.ncu-rep out.md [title]"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'sm__icc_request_hit_rate.pct', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_warps_issue_stalled_mio_throttle',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'smsp__pcsamp_warps_issue_stalled_no_instructions']
lines = ["# " + title, "", "source: `%s` (ncu --set full --clock-control none --import-source on)" % rep, ""]
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if name in seen:
        continue
    seen.add(name)
    lines += ["## " + name, "", "| metric | unit | value |", "|---|---|---|"]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("| %s | %s | %s |" % (k, units[i], r[i]))
    lines.append("")
reg = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_regions.py"), rep], capture_output=True, text=True).stdout
# keep the first block per kernel
blocks, cur, names = [], None, set()
for ln in reg.splitlines():
    if not ln.startswith("  "):
        nm = ln.split("|")[0]
        cur = [] if nm not in names else None
        names.add(nm)
        if cur is not None:
            blocks.append(cur)
    if cur is not None:
        cur.append(ln)
lines += ["## SASS regions (share of executed warp-instructions / stall samples, avg active threads)", "", "```"]
for b in blocks:
    lines += b
lines += ["```", ""]
open(out, "w").write("\n".join(lines))
print("wrote", out)
