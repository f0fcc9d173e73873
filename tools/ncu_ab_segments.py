import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
names=sys.argv[2].split(',')
by=collections.OrderedDict()
for r in rows: by.setdefault(r[0],{})[r[12]]=float(r[14].replace(',',''))
L=list(by.values())
per=len(L)//(2*len(names))
for i,nm in enumerate(names*2):
    seg=L[i*per+10:(i+1)*per]
    t=sum(d['gpu__time_duration.sum'] for d in seg)/len(seg); c=sum(d['sm__cycles_elapsed.max'] for d in seg)/len(seg)
    tp=sum(d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'] for d in seg)/len(seg)
    print(f"{nm:12s} {t/1e6:.3f} ms {c/1e6:.3f} Mcyc {c/t:.3f} GHz tensor {tp:.1f}%")
