import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if True: print(d['op'],d['n'],d['kernel'],round(d['ms_best'],3),round(d['TFLOPs'],2))
