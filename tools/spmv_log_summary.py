"""Group the per-call SpMV log (AMGB_SPMV_LOG=1, stderr of a setup) by matrix.

    AMGB_SPMV_LOG=1 python tools/profile_setup.py poisson7 128 1 2> a.log
    python tools/spmv_log_summary.py a.log [b.log ...]        # one column pair per log

Lines: ``spmv R x C nnz N  T ms  B GB/s``; only the LAST setup of each log is used (the calls
after the last marker line ``profiled setup`` do not exist, so the log is cut in halves by count).
"""
import collections
import re
import sys


def load(path, setups):
    calls = []
    for line in open(path, errors="replace"):
        m = re.match(r"spmv (\d+) x (\d+) nnz (\d+)\s+([\d.]+) ms", line)
        if m:
            calls.append((int(m.group(1)), int(m.group(2)), int(m.group(3)), float(m.group(4))))
    per = len(calls) // setups
    calls = calls[-per:]
    agg = collections.OrderedDict()
    for r, c, nnz, ms in calls:
        k = (r, c, nnz)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    return agg


def main():
    setups = 2
    logs = [load(p, setups) for p in sys.argv[1:]]
    keys = sorted(logs[0], key=lambda k: -logs[0][k][1])
    print("# totals: " + ", ".join("%s %.1f ms" % (p, sum(v[1] for v in l.values())) for p, l in zip(sys.argv[1:], logs)))
    print("%-40s %6s %6s | %s" % ("rows x cols nnz", "row", "calls", " | ".join(sys.argv[1:])))
    for k in keys[:60]:
        r, c, nnz = k
        cells = []
        for l in logs:
            if k in l:
                n, ms = l[k]
                cells.append("%8.1f ms %6.0f GB/s" % (ms, (12.0 * nnz + 16.0 * r) * n / (ms * 1e-3) / 1e9))
            else:
                cells.append("-")
        print("%-40s %6d %6d | %s" % ("%d x %d %d" % k, nnz // max(r, 1), logs[0][k][0], " | ".join(cells)))


if __name__ == "__main__":
    main()
