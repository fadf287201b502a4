"""Print the headline numbers of bench.py JSON lines found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    for l in open(f):
        l = l.strip()
        if l.startswith('{'):
            d = json.loads(l)
            print(f, d.get('n_gpus'), 'value', round(d['value']), 'e2e', round(d['e2e']['value']),
                  'serial', round(d.get('serial', {}).get('value', 0)),
                  {k: round(v, 3) for (k, v) in d.get('stage_ms_per_step', {}).items()},
                  'frac', round(d.get('roofline', {}).get('frac', 0), 4), d.get('clocks'))
