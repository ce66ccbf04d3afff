"""Pick the judged metrics out of an ``ncu --set full`` report (read here, no GPU needed).

    python benchmarks/ncu_summary.py gpurun_out/paint3.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
    'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
    'smsp__inst_executed_op_global_red.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
    'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed',
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel: %s" % d.get('Kernel Name'))
        for m in WANT:
            if m in d:
                print("%-72s %s %s" % (m, d[m], units[hdr.index(m)]))
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                         if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio')
                         and v not in ('', 'n/a')), reverse=True)[:6]
        print("top warp stall reasons (warps per issue-active cycle):")
        for v, k in stalls:
            print("  %-70s %.2f" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
        print()


if __name__ == '__main__':
    main()
