"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys

def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v = v / 1000.0 if unit in ('ns', 'nsecond') else (v * 1000.0 if unit in ('ms', 'msecond') else v)
        k = row['Kernel Name'].replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('lg::', '')
        k = re.sub(r'\(.*$', '', k)[:100]
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2] = max(agg[k][2], v)
        tot += v
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
    print("%10s %6s %5s %9s %9s  kernel" % ("sum_us", "share", "n", "avg_us", "max_us"))
    for k, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%10.1f %5.1f%% %5d %9.1f %9.1f  %s" % (t, 100 * t / tot, n, t / n, mx, k))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
