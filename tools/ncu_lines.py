"""Per-source-line stall samples / instructions of one kernel in an ncu report (compiled with -lineinfo):
python tools/ncu_lines.py report.ncu-rep kernel_name [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fil = ''
out = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fil = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if r[0] == 'Function Name' or hdr is None:
        continue
    if r[0] != '' and r[2] == '-':
        try:
            out.append((int(r[4]), int(r[7]), fil, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print('total samples', tot, 'total warp instructions', toti)
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * s / tot, 100.0 * i / toti, f, ln, src))
