"""One line per kernel of an `ncu --page raw --csv` export: duration, issue / pipe utilisation, stall reasons (warps per issue)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    def f(k):
        try: return float(d[k].replace(',', ''))
        except Exception: return -1.0
    st = lambda n: f('smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % n)
    print(d['Kernel Name'][:40], 'ms %.2f' % f('gpu__time_duration.sum'),
          'issue %.1f%% alu %.1f%% fmaheavy %.1f%% lsu %.1f%%' % (f('smsp__issue_active.avg.pct_of_peak_sustained_active'), f('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'),
                                                     f('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'), f('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active')),
          'active/elapsed %.2f' % (f('sm__cycles_active.avg') / max(f('sm__cycles_elapsed.avg'), 1)),
          'stalls: noinst %.2f notsel %.2f math %.2f wait %.2f longsb %.2f shortsb %.2f dispatch %.2f branch %.2f' % tuple(
              st(n) for n in ('no_instruction', 'not_selected', 'math_pipe_throttle', 'wait', 'long_scoreboard', 'short_scoreboard', 'dispatch_stall', 'branch_resolving')),
          'inst %.3g' % f('smsp__inst_executed.sum'),
          'dram R+W %.2f GB' % ((f('dram__bytes_read.sum') + f('dram__bytes_write.sum')) / 1e9 if 'dram__bytes_read.sum' in d else -1))
