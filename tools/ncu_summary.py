"""Summarise an ncu report (--set full) as a markdown table per kernel launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title / command line" > profiles/<name>.md

Run where `ncu` is installed (the builder container reads reports without a GPU)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid size (CTAs)'),
    ('launch__block_size', 'threads per CTA'),
    ('launch__registers_per_thread', 'registers per thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic shared memory per CTA'),
    ('launch__occupancy_limit_registers', 'CTAs/SM allowed by registers'),
    ('launch__occupancy_limit_shared_mem', 'CTAs/SM allowed by shared memory'),
    ('launch__waves_per_multiprocessor', 'waves per SM'),
    ('dram__bytes_read.sum', 'DRAM bytes read'),
    ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput, % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput, % of peak'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy'),
    ('sm__inst_executed.avg.per_cycle_active', 'warp instructions per cycle per SM (max 4)'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active'),
    ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'ALU pipe active'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
    ('smsp__warps_active.avg.per_cycle_active', 'warps resident per scheduler'),
    ('smsp__warps_eligible.avg.per_cycle_active', 'warps eligible per scheduler per cycle'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else '')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print('# %s\n' % (title or rep))
    print('Source report: `%s` (kept in gpurun_out/, not committed).  Numbers under ncu are cold-cache and serialised; '
          'they are evidence for *where* the time goes, never bench values.\n' % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('## `%s`\n' % d.get('Kernel Name', '?'))
        print('| metric | value |\n|---|---|')
        for key, label in METRICS:
            if key in d:
                print('| %s (`%s`) | %s %s |' % (label, key, d[key], u.get(key, '')))
        st = {}
        for h, v in d.items():
            if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h:
                try:
                    st[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(v)
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        top = sorted(st.items(), key=lambda kv: -kv[1])[:10]
        print('\nWarp-state sampling: ' + ', '.join('%s %.1f %%' % (k, 100 * v / tot) for k, v in top) + '\n')


if __name__ == '__main__':
    main()
