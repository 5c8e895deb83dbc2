import csv, subprocess, sys
rep=sys.argv[1]; kern=sys.argv[2]; topn=int(sys.argv[3]) if len(sys.argv)>3 else 25
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
# may contain several kernels: split at "Function Name"
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Function Name':
        cur={'name':r[1],'rows':[]}; blocks.append(cur)
    elif cur is not None: cur['rows'].append(r)
for b in blocks:
    want=sys.argv[4] if len(sys.argv)>4 else ''
    if want and want not in b['name']: continue
    rs=b['rows']
    h=[i for i,r in enumerate(rs) if r and r[0]=='Line No'][0]
    hdr=rs[h]; ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples'); ti=hdr.index('Thread Instructions Executed')
    agg=[]
    for r in rs[h+1:]:
        if r and r[0].isdigit():
            try: agg.append((int(r[0]), r[1], int(r[ci]), int(r[si]), int(r[ti])))
            except: pass
    tot=sum(a[2] for a in agg); tots=sum(a[3] for a in agg)
    print(b['name'], "total inst", tot, "samples", tots)
    for a in sorted(agg,key=lambda x:-x[3])[:topn]:
        print(f"{a[0]:4d} inst {a[2]/max(tot,1)*100:5.1f}% samp {a[3]/max(tots,1)*100:5.1f}% thr/inst {a[4]/max(a[2],1):5.1f} | {a[1][:105]}")
