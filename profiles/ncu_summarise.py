import csv, sys, subprocess, collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__cycles_active.avg','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','launch__grid_size','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum']
for w in want:
    if w in hdr:
        i=hdr.index(w); print(f'{w:64s} {units[i]:14s}', [r[i] for r in rows[2:]])
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
secs=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
i0=secs[0]; i1=secs[1] if len(secs)>1 else len(rows)
hdr=rows[i0+1]; body=[r for r in rows[i0+2:i1] if len(r)==len(hdr)]
H={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[H['# Samples']]) for r in body)
print('SASS instrs', len(body), 'samples', tot, 'warp-instr executed', sum(int(r[H['Instructions Executed']]) for r in body))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={s:sum(int(r[H[s]]) for r in body) for s in stalls}
for s,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]: print(f'  {s:28s}{v:7d} {100*v/max(tot,1):5.1f}%')
top=sorted(range(len(body)), key=lambda k:-int(body[k][H['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 14]
for k in sorted(top):
    r=body[k]; st=sorted(((int(r[H[s]]),s) for s in stalls), reverse=True)[:1]
    print('  ',k, r[H['Source']][:64].ljust(66), r[H['# Samples']].rjust(5), st)
ops=collections.Counter()
for r in body:
    n=int(r[H['Instructions Executed']]); op=r[H['Source']].split()[0] if not r[H['Source']].strip().startswith('@') else r[H['Source']].split()[1]
    ops[op.split('.')[0]]+=n
nw=max(int(r[H['Instructions Executed']]) for r in body[:3])
print('  per-warp opcode mix:', [(k,round(v/nw,1)) for k,v in ops.most_common(22)])
