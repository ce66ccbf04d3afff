"""Summarise an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv``) of
``bench.py --steps 2 --warmup W``: per-kernel totals of ONE timed step, i.e. the launches
between the two L2-flush fills (256 MiB ``fill_``) that precede the two timed steps.

    python benchmarks/summarize_launches.py gpurun_out/launches.csv > profiles/<name>_step_summary.txt
"""
import collections
import csv
import re
import sys


def read(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
        rows.append((int(r['ID']), r['Kernel Name'], v * scale))
    return rows


def short(name):
    name = re.sub(r'\(.*', '', name).replace('void ', '').replace('vpm::', '')
    return name[:60]


def main():
    rows = read(sys.argv[1])
    fills = [i for i, (_, n, us) in enumerate(rows) if 'FillFunctor' in n and us > 20.0]
    if len(fills) < 2:
        raise SystemExit("could not find two L2-flush fills (found %d)" % len(fills))
    # the first pair of big fills with a whole step of launches between them brackets one step
    # (other big fills: zero-initialised meshes while the problem is built)
    pair = next(((fills[i] + 1, fills[i + 1]) for i in range(len(fills) - 1) if fills[i + 1] - fills[i] > 100), None)
    if pair is None:
        raise SystemExit("no two big fills with a step between them")
    a, b = pair
    step = rows[a:b]
    tot = sum(us for _, _, us in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, n, us in step:
        e = agg[short(n)]
        e[0] += 1
        e[1] += us
    print("one timed step = launches %d..%d of %d; per-launch times are cold-cache and serialised: compare SHARES"
          % (rows[a][0], rows[b - 1][0], len(rows)))
    print("total %.2f ms over %d launches" % (tot / 1e3, len(step)))
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s n=%4d %9.3f ms %6.1f%%  avg %7.1f us" % (k, n, us / 1e3, 100 * us / tot, us / n))


if __name__ == '__main__':
    main()
