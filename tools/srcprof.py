"""Summarise `ncu --page source --csv --print-source cuda,sass` output: stall samples per CUDA source line.
usage: python scratch/srcprof.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if '# Samples' in r]
h = rows[hi[0]]
S = h.index('# Samples')
end = hi[1] if len(hi) > 1 else len(rows)
lines = []
for r in rows[hi[0] + 1:end]:
    if len(r) != len(h) or not r[0].strip().isdigit():
        continue
    try:
        lines.append((int(r[S]), int(r[0]), r[1]))
    except ValueError:
        pass
tot = sum(x[0] for x in lines)
print('total samples', tot)
for n, ln, src in sorted(lines, reverse=True)[:top]:
    print('%7d %5.1f%%  L%-5d %s' % (n, 100.0 * n / tot, ln, src.strip()[:120]))
