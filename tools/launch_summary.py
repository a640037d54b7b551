"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[hdr + 2:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        u = r[ui]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        k = r[ki]
        k = k[:k.index('(')] if '(' in k else k
        if 'spin_kernel' in k:  # bench.py's queue-filling spin in front of its event-bracketed eager steps
            continue
        agg[k[:90]][0] += 1
        agg[k[:90]][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print("launches %d, total %.1f us" % (n, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s n=%4d  %10.1f us  %5.1f%%  avg %7.1f us" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    main()
