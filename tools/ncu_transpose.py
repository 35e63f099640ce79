"""`ncu -i X.ncu-rep --page raw --csv` (one row per launch, ~2400 columns) -> one line per metric per captured launch:
    metric,unit,value,launchN          (the layout of profiles/*_ncu_full.csv; bench.py reads dram__bytes_* from it)
usage: python tools/ncu_transpose.py raw.csv out.csv "header comment" ["second header comment"]"""
import csv, sys
raw, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H, U = rows[hdr], rows[hdr + 1]
launches = [r for r in rows[hdr + 2:] if len(r) == len(H)]
skip = {'ID', 'Process ID', 'Process Name', 'Host Name', 'Context', 'Stream', 'Device', 'CC', 'Section Name', 'Metric Name', 'Metric Unit'}
with open(out, 'w') as f:
    for c in sys.argv[3:]:
        f.write('# %s\n' % c)
    for n, r in enumerate(launches):
        for h, u, v in zip(H, U, r):
            if h in skip:
                continue
            f.write('%s,%s,%s,launch%d\n' % (h, u, v.replace(',', ';'), n))
print(out, len(launches), 'launches')
