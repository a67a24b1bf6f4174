import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['op'], d['n'], d.get('variant'), d['kernel'], round(d['ms_best'], 3))
