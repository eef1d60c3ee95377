"""Top SASS instructions by stall samples for launch #k of an .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys
rep, k, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Kernel Name': cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(row)
b = blocks[k]; hdr = b[0]; rows = b[1:]
si = hdr.index('# Samples'); src = hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in rows)
print('total samples', tot, 'instructions', len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i][si]))[:top]
for i in sorted(idx):
    r = rows[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  {r[src].strip()[:90]:90s} {st}")
