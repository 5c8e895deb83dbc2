import csv, subprocess, sys
rep=sys.argv[1]; kern=sys.argv[2]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','lts__t_bytes.sum','l1tex__t_bytes.sum','launch__occupancy_limit_registers','launch__waves_per_multiprocessor']
want+= [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    if kern in r[idx['Kernel Name']]:
        print('----', r[idx['Kernel Name']][:60])
        for w in want:
            if w in idx:
                try: v=float(r[idx[w]].replace(',',''))
                except: v=r[idx[w]]
                if isinstance(v,float) and 'stalled' in w and v<0.15: continue
                print(f"  {w.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','')} = {r[idx[w]]} {units[idx[w]]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern.replace('<','.').replace('>','.').replace('<','.').replace('>','.')],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=[i for i,r in enumerate(rows) if r and r[0]=='Line No'][0]
hdr=rows[h]; ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples'); ti=hdr.index('Thread Instructions Executed')
agg=[]
for r in rows[h+1:]:
    if r and r[0].isdigit():
        try: agg.append((int(r[0]), r[1], int(r[ci]), int(r[si]), int(r[ti])))
        except: pass
tot=sum(a[2] for a in agg); tots=sum(a[3] for a in agg)
print("total inst", tot, "samples", tots)
for a in sorted(agg,key=lambda x:-x[3])[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(f"{a[0]:4d} inst {a[2]/tot*100:5.1f}% samp {a[3]/tots*100:5.1f}% thr/inst {a[4]/max(a[2],1):5.1f} | {a[1][:105]}")
