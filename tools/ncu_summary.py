import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','launch__block_size','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__waves_per_multiprocessor','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','sm__inst_executed_pipe_fp64.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx: print(' ', w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''), '=', r[idx[w]], units[idx[w]])

# --traffic OUT.json TAG: record dram__bytes_read + dram__bytes_write per launch of every kernel in the capture under
# the key "<kernel>|<TAG>" (TAG = "C<config>|<n_lnc>|<fitter>", what bench.py's ncu_traffic() looks up); the capture's
# file name is kept beside the number so a stale entry can be traced.
if '--traffic' in sys.argv:
    import json, os, re
    out, tag = sys.argv[sys.argv.index('--traffic') + 1], sys.argv[sys.argv.index('--traffic') + 2]
    table = json.load(open(out)) if os.path.exists(out) else {}
    to_bytes = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    acc = {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        m = re.match(r'(?:void )?(\w+)(<[^>]*>)?', name)
        base = m.group(1)
        targs = (m.group(2) or '')
        if base == 'k_front' and targs:               # k_front<double, 0> -> k_front<double>
            targs = '<' + targs[1:-1].split(',')[0].strip() + '>'
        key = base + (targs if base == 'k_front' else '')
        b = sum(float(r[idx[k]]) * to_bytes[units[idx[k]]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        acc.setdefault(key, []).append((b, float(r[idx['gpu__time_duration.sum']])))
    for key, v in acc.items():
        table[f'{key}|{tag}'] = {'dram_bytes': sum(x[0] for x in v) / len(v), 'launches': len(v),
                                 'gpu_time_us_under_ncu': sum(x[1] for x in v) / len(v), 'capture': os.path.basename(sys.argv[1])}
    json.dump(table, open(out, 'w'), indent=1, sort_keys=True)
