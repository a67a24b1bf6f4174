import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['op']=='Dpotrf' and d['n']==32: print(d['op'],d['n'],d['variant'],d['kernel'],round(d['ms_best'],3),round(d['ms_mean'],3),round(d['frac'],3))
