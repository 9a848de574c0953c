import csv,sys,gzip,io
rows=list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
hi=next(i for i,r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h=rows[hi]; isrc=h.index('Source'); iex=h.index('Instructions Executed'); ismp=h.index('# Samples')
st=[(i,n) for i,n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
body=rows[hi+1:]
tot=sum(int(r[ismp] or 0) for r in body if len(r)>ismp)
idx=sorted(range(len(body)), key=lambda i:-int(body[i][ismp] or 0))[:int(sys.argv[2])]
for i in idx:
    r=body[i]
    reasons=sorted(((int(r[j] or 0),n) for j,n in st), reverse=True)[:2]
    print(i, f"{int(r[ismp])/tot*100:5.2f}%", r[iex], r[isrc].strip()[:60], reasons)
