"""Prints a gpurun_out/*lab*.log of tools/kl_lab.py as a table."""
import json
import sys

for l in open(sys.argv[1]):
    l = l.strip()
    if not l.startswith('{'):
        print(l[:200])
        continue
    d = json.loads(l)
    if 'accuracy_vs_fp64' in d:
        print('%-10s acc %s' % (d['variant'], {k: '%.1e' % v for k, v in d['accuracy_vs_fp64'].items()}))
    elif 'sustained_ms_per_pass_median' in d:
        print('%-10s sustained %s x%d: %.3f ms per pass (min %.3f max %.3f)' % (d['variant'], '+'.join(d['ops']), d['reps'], d['sustained_ms_per_pass_median'], d['min'], d['max']))
    elif 'reference' in d:
        print('reference %s median %.3f min %.3f max %.3f ms' % (d['reference'], d['median_ms'], d['min_ms'], d['max_ms']))
    else:
        print('%-10s dbg=%-6s ' % (d['variant'], hex(d['dbg'])) + '  '.join(
            '%s %.2f/%.2f r=%.3f' % (k, d[k]['median_ms'], d[k]['min_ms'], d[k].get('ratio_to_ref_median', 0)) for k in ('ah', 'wta', 'kl_uht', 'kl_wtu') if k in d))
