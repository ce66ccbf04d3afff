"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every hand-written
entry point, from an ``ncu --set full`` capture of ``benchmarks/ncu_targets.py <nmesh>`` ->
``profiles/roofline_traffic.json`` (read by bench.py for ``roofline.traffic``).

    python benchmarks/make_traffic.py gpurun_out/r2_force_512.ncu-rep 512
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel-name fragment -> entry point (an entry point sums its kernels; memsets are not kernels
# of ours and add their 8 B/cell on top)
MAP = [('k_paint_agg3', 'paint_cols3'), ('k_paint_agg<', 'paint'), ('k_readout_grad4', 'readout_grad4'),
       ('k_readout3', 'readout3'), ('k_take_rows3', 'take_rows'), ('k_fd4_grad3', 'fd4_neg_gradient'),
       ('k_fd4_adjoint', 'fd4_neg_gradient_adjoint'), ('k_drift_pos', 'drift_pos'), ('k_transfer', 'transfer'),
       ('k_keys_count', 'layout_build'), ('k_scan_', 'layout_build'), ('k_scatter_ids', 'layout_build'),
       ('k_fix_', 'layout_build'), ('k_copy_big', 'layout_build')]


SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def launches_from_report(rep):
    """(kernel name, dram bytes) per launch of an .ncu-rep"""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    for r in rows[2:]:
        yield r[ik], (float(r[ir].replace(',', '')) * SCALE[units[ir]] + float(r[iw].replace(',', '')) * SCALE[units[iw]])


def launches_from_csv(path):
    """the same from ``ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file`` (one
    row per launch and metric)"""
    lines = [l for l in open(path, newline='') if not l.startswith('==')]
    acc = {}
    order = []
    for r in csv.DictReader(lines):
        if r.get('Metric Name') not in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            continue
        k = int(r['ID'])
        if k not in acc:
            acc[k] = [r['Kernel Name'], 0.0]
            order.append(k)
        acc[k][1] += float(r['Metric Value'].replace(',', '')) * SCALE[r['Metric Unit']]
    for k in order:
        yield acc[k][0], acc[k][1]


def main():
    rep, nmesh = sys.argv[1], sys.argv[2]
    what = sys.argv[3] if len(sys.argv) > 3 else "benchmarks/ncu_targets.py %s" % nmesh
    per = {}
    calls = {}
    primary = {}
    for frag, e in MAP:
        primary.setdefault(e, frag)         # launches of the first-listed kernel count the calls
    for name, b in (launches_from_csv(rep) if rep.endswith('.csv') else launches_from_report(rep)):
        ep = next((e for frag, e in MAP if frag in name), None)
        if ep is None:
            continue
        per[ep] = per.get(ep, 0.0) + b
        if primary[ep] in name:
            calls[ep] = calls.get(ep, 0) + 1
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    try:
        d = json.load(open(path))
        if not isinstance(d.get('source'), dict):
            d = {}
    except Exception:
        d = {}
    d[str(nmesh)] = {ep: v / max(calls.get(ep, 1), 1) for ep, v in sorted(per.items())}
    d.setdefault('source', {})[str(nmesh)] = (
        "ncu capture of %s (%s): dram__bytes_read.sum + dram__bytes_write.sum, mean per call of the entry point "
        "over the captured launches (zero fills by cudaMemset excluded)" % (what, os.path.basename(rep)))
    json.dump(d, open(path, 'w'), indent=1)
    print(json.dumps(d[str(nmesh)], indent=1))


if __name__ == '__main__':
    main()
