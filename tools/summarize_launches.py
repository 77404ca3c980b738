"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/profile_step.py:
prefill chunk = launches before the last 2*n_dec of ours, decode step = the last n_dec (228: 7 per layer + 4)."""
import csv, collections, re, sys
path = sys.argv[1]; n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 228
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(int(r['ID']), r['Kernel Name'], float(r['Metric Value']), r['Grid Size']) for r in csv.DictReader(lines)
        if r.get('Metric Name') == 'gpu__time_duration.sum']
def short(n):
    n = n.replace('<unnamed>::', '').replace('unnamed>::', '')
    m = re.match(r'(void )?(rr::)?(\w+)(<[^>]*>)?', n); return (m.group(3) + (m.group(4) or '')) if m else n[:40]
print(f"{path}: {len(rows)} launches (per-launch times are cold-cache + serialised: compare SHARES)")
for name, part in (('PREFILL chunk (16 x 512 tokens)', rows[:len(rows) - 2 * n_dec]), ('DECODE step (64 rows, ctx 577)', rows[-n_dec:])):
    if not part: continue
    agg = collections.OrderedDict()
    for _, k, v, g in part:
        a = agg.setdefault(short(k) + ' grid=' + g, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, _, v, _ in part)
    print('%s: total %.3f ms over %d launches' % (name, tot / 1e6, len(part)))
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('  %-62s n=%4d  %9.1f us  %5.1f%%  avg %6.1f us' % (k[:62], c, v / 1e3, 100 * v / tot, v / 1e3 / c))
