import csv,sys
from collections import defaultdict
rows=list(csv.reader(open(sys.argv[1])))
start=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
cols=rows[start]; ki=cols.index("Kernel Name"); vi=cols.index("Metric Value")
agg=defaultdict(list)
for r in rows[start+2:]:
    if len(r)>vi:
        try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
        except: pass
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1))
