"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, re, collections, sys
def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]; data = rows[hi + 1:]
    kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mu = hdr.index('Metric Unit')
    tot = collections.Counter(); cnt = collections.Counter(); mx = collections.Counter()
    for r in data:
        if len(r) <= mv: continue
        name = r[kn]
        t = float(r[mv].replace(',', ''))
        u = r[mu]
        t *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(u, 1)
        if 'k_parallel_for' in name:
            m2 = re.findall(r'amgb::([A-Za-z_0-9]+)\(', name)
            key = 'pf:' + (m2[-1] if m2 else name[:50])
        else:
            key = re.sub(r'^void ', '', name).split('(')[0][:48]
        tot[key] += t; cnt[key] += 1; mx[key] = max(mx[key], t)
    T = sum(tot.values())
    print("total kernel time %.3f ms over %d launches" % (T / 1e6, sum(cnt.values())))
    for k, v in tot.most_common(top):
        print("%-50s %9.3f ms %5.1f%%  n=%5d  max=%9.3f ms" % (k, v / 1e6, 100 * v / T, cnt[k], mx[k] / 1e6))
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
