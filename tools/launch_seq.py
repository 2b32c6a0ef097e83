"""Print the kernel sequence of an ncu --metrics gpu__time_duration.sum --csv log compactly."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
seq = []
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    k = r['Kernel Name'].split('(')[0].replace('pcreg::', '').replace('void ', '')
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
    seq.append((k, ms))
skip = ('k_grid', 'k_scan', 'k_build', 'at::', 'k_transpose', 'k_src')
print(' '.join('%s:%.3f' % (k[:16], ms) for k, ms in seq if not k.startswith(skip)))
