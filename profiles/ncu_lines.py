import csv,re,sys,subprocess
from collections import Counter
rep, kern_mangled, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv)>4 else 40
warpsteps = float(sys.argv[5]) if len(sys.argv)>5 else 262144.0
subprocess.run(f"ncu -i {rep} --page source --csv 2>/dev/null > /tmp/_src.csv", shell=True)
subprocess.run(f"ncu -i {rep} --page raw --csv 2>/dev/null > /tmp/_raw.csv", shell=True)
dis = subprocess.run(f"nvdisasm -g -c {cubin} 2>/dev/null", shell=True, capture_output=True, text=True).stdout.splitlines()
start = [i for i,l in enumerate(dis) if l.startswith(kern_mangled+':')][0]
addr_line={}; cur=None
for l in dis[start:]:
    if l.startswith('//---------------------') and addr_line: break
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*);',l)
    if m: addr_line[int(m.group(1),16)]=(cur,m.group(2))
rows=list(csv.reader(open('/tmp/_src.csv')))
hdr_i=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hdr_i[0]]; end=hdr_i[1]-1 if len(hdr_i)>1 else len(rows)
body=rows[hdr_i[0]+1:end]
ie=h.index('Instructions Executed'); smp=h.index('# Samples'); src=h.index('Source')
base=int(body[0][0],16)
byline=Counter(); samp=Counter(); ops=Counter()
for r in body:
    if not r[ie].isdigit(): continue
    a=int(r[0],16)-base
    loc=addr_line.get(a,(None,''))[0]
    byline[loc]+=int(r[ie]); samp[loc]+=int(r[smp]) if r[smp].isdigit() else 0
    t=r[src].split(); op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op]+=int(r[ie])
tot=sum(byline.values())
print('total warp-inst', tot, 'per warp-step', tot/warpsteps)
for loc,n in byline.most_common(top):
    print(f'{str(loc):34s} {n/warpsteps:7.1f}/ws {100*n/tot:5.1f}% smp {samp[loc]}')
print(' '.join(f'{o}:{n/warpsteps:.1f}' for o,n in ops.most_common(30)))
raw=list(csv.reader(open('/tmp/_raw.csv'))); hh=raw[0]
for k in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum']:
    if k in hh: print(k, raw[2][hh.index(k)])
for k in hh:
    if 'warp_issue_stalled' in k and k.endswith('_per_warp_active.pct'):
        v=float(raw[2][hh.index(k)])
        if v>3: print(k.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''), round(v,1))
