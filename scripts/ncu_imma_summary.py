"""ncu report of the IMMA kernels -> the text summary kept under profiles/ (launch metrics, stall reasons per issue, instruction mix per row):
    python scripts/ncu_imma_summary.py gpurun_out/r2_imma_N677.ncu-rep [rows] > profiles/r2_imma_N677_summary.txt"""
import csv,sys,subprocess,collections,re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:70])
    for k in want:
        if k in hdr: print('   ',k,r[hdr.index(k)])
    st={k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):float(r[hdr.index(k)]) for k in hdr if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')}
    print('    stalls/issue:',', '.join(f'{k} {v:.2f}' for k,v in sorted(st.items(),key=lambda x:-x[1])[:8]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
kern=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]};kern.append(cur);continue
    if r and r[0]=='Address': cur['hdr']=r;continue
    if cur is not None and r: cur['rows'].append(r)
seen=set()
nrows=int(sys.argv[2]) if len(sys.argv)>2 else 262144
for k in kern:
    if k['name'] in seen: continue
    seen.add(k['name'])
    h=k['hdr'];iS=h.index('Source');iE=h.index('Instructions Executed');iSm=h.index('# Samples')
    tot=sum(int(r[iE]) for r in k['rows']);ts=sum(int(r[iSm]) for r in k['rows'])
    per=collections.Counter();pers=collections.Counter()
    for r in k['rows']:
        t=r[iS].split(); op=t[1] if t[0].startswith('@') else t[0]; op=op.split('.')[0]
        per[op]+=int(r[iE]);pers[op]+=int(r[iSm])
    print(k['name'][:70],'inst/row',tot/nrows)
    print('   ',' '.join(f'{op}:{c/nrows:.0f}({pers[op]/ts*100:.0f}%)' for op,c in per.most_common(14)))
