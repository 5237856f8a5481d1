"""Stall picture of a kernel from an `ncu --set full --import-source on` report (development tool).

    python tools/ncu_stalls.py gpurun_out/r17_f2p.ncu-rep [more.ncu-rep ...]

Prints, for the first kernel in each report: duration, instructions, issue-slot utilisation, DRAM rate,
registers / occupancy limits, pipe utilisation, the warp-stall reasons per issued instruction, the SASS
instructions with the most stall samples (with their execution counts), and the execution-count buckets
(which show the hot loop's length: the bucket whose count equals warps x rows).  Also writes the annotated
SASS listing (index, address, executions, samples, instruction) to /tmp/<report>_sass.txt.
"""
import csv, subprocess, sys
def raw(rep):
    out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    return dict(zip(rows[0],rows[2]))
def src(rep):
    out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    secs=[];cur=None
    for r in rows:
        if r and r[0]=="Kernel Name": cur={'name':r[1],'rows':[]}; secs.append(cur)
        elif r and r[0]=="Address": cur['hdr']=r
        elif cur is not None and r: cur['rows'].append(r)
    return secs[0]
for rep in sys.argv[1:]:
    d=raw(rep)
    print('=====',rep)
    for k in ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes.sum.per_second','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active']:
        if k in d: print(' ',k,d[k])
    st={h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):float(d[h]) for h in d if 'stalled' in h and 'per_issue_active' in h and 'not_issued' not in h}
    print('  stalls:',' '.join('%s=%.2f'%(k,v) for k,v in sorted(st.items(),key=lambda kv:-kv[1]) if v>0.05))
    s=src(rep); h=s['hdr']; ix={n:i for i,n in enumerate(h)}
    tot=sum(int(r[ix['# Samples']]) for r in s['rows'])
    rs=sorted(s['rows'],key=lambda r:-int(r[ix['# Samples']] or 0))[:14]
    for r in rs: print('   %5.1f%%'%(100*int(r[ix['# Samples']])/tot), r[1][:70], r[ix['Instructions Executed']])
    from collections import Counter
    c=Counter()
    for r in s['rows']: c[int(r[ix['Instructions Executed']])]+=1
    print('  exec-count buckets:',[(k,v) for k,v in sorted(c.items(), key=lambda kv:-kv[0]*kv[1])[:6]])
    tag=rep.split('/')[-1].replace('.ncu-rep','')
    with open('/tmp/%s_sass.txt'%tag,'w') as f:
        for i,r in enumerate(s['rows']):
            f.write('%5d %s %9s %6s  %s\n'%(i, r[0][-5:], r[ix['Instructions Executed']], r[ix['# Samples']], r[1]))
