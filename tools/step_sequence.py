"""Print the kernel sequence of the LAST training step in an ncu launch list (steps are delimited by opt_multi_kernel)."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    gi = h.index('Grid Size') if 'Grid Size' in h else None
    out = []
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
        if 'spin_kernel' in k:
            continue
        out.append((k.replace('void ', '').replace('dk::', ''), v, r[gi] if gi is not None else ''))
    return out


def main():
    seq = load(sys.argv[1])
    ends = [i for i, (k, _, _) in enumerate(seq) if k.startswith('opt_multi_kernel')]
    a, b = ends[-2] + 1, ends[-1] + 1
    tot = 0.0
    for k, v, g in seq[a:b]:
        tot += v
        print("%-60s %8.1f us  grid %s" % (k[:60], v, g))
    print("step total %.1f us, %d launches" % (tot, b - a))


if __name__ == "__main__":
    main()
