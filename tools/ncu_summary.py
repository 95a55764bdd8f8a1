"""Summarise an .ncu-rep: key raw metrics per kernel + opcode histogram + top stall lines (development tool)."""
import csv, re, subprocess, sys, io
from collections import defaultdict
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warp_latency_per_inst_issued.ratio','sm__cycles_elapsed.max','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum','sm__inst_executed_pipe_fp64.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_op_shared_ld.sum']
for r in rows[2:]:
    print('-----')
    for w in want:
        for i, h in enumerate(hdr):
            if h == w: print(f'{w:90s} {r[i][:90]} {units[i]}')
args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if kern: args += ["--kernel-name", "regex:" + kern]
src = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if r and r[0] == 'Address': h = i; break
hdr = rows[h]
si, ei, sm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
op = defaultdict(int); samp = defaultdict(int); tot = stot = 0
lines = []
for r in rows[h+1:]:
    if len(r) <= sm or r[0] == 'Address' or not r[ei].isdigit(): continue
    s = r[si].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s)
    o = m.group(2).split('.')[0] if m else s
    n = int(r[ei]); op[o] += n; tot += n
    samp[o] += int(r[sm]); stot += int(r[sm])
    lines.append((int(r[sm]), n, r[0], s))
print('total warp-inst', tot, 'samples', stot)
for k in sorted(op, key=op.get, reverse=True)[:22]:
    print(f'{k:10s} {op[k]:12d} {100*op[k]/tot:5.1f}%  stall-samples {100*samp[k]/max(stot,1):5.1f}%')
