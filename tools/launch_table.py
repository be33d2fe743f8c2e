"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum --csv` launch list (gpurun_out/*_launches.csv)."""
import collections, csv, re, sys

def table(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
        agg[(name, row['Grid Size'], row['Block Size'])].append(v)
    tot = sum(sum(v) for v in agg.values())
    out = []
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append((k[0], k[1], k[2], len(v), sum(v) / len(v), sum(v) / tot))
    return out

if __name__ == '__main__':
    md = '--md' in sys.argv
    for name, grid, block, n, avg, share in table(sys.argv[1]):
        if md:
            print('| %s | %s x %s | %d | %.1f | %.3f |' % (name, grid, block, n, avg, share))
        else:
            print('%-42s %18s %14s n=%3d avg=%8.1f us share=%.3f' % (name[:42], grid, block, n, avg, share))
