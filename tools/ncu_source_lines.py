"""Stall samples of an ncu report by SOURCE LINE, without the source tree on the GPU box: joins ncu's SASS page (per
instruction samples and stall reasons, in address order) with `nvdisasm -g` of the same cubin (per instruction file:line
from -lineinfo); the two list the kernel's instructions in the same order.
usage: ncu_source_lines.py report.ncu-rep kernel.cubin KERNEL_NAME_SUBSTRING [top_n]
  (cubin: `cuobjdump -xelf all sgmcmcjax_b200/libsgmcmc_b200.so` of the build that was profiled)"""
import csv,re,subprocess,sys,collections
rep,cubin,kern=sys.argv[1],sys.argv[2],sys.argv[3]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; data=rows[2:]
ci={n:i for i,n in enumerate(hdr)}
dis=subprocess.run(['nvdisasm','-g','-c',cubin],capture_output=True,text=True).stdout.splitlines()
ins=[];fn=None;fl=None;ln=None;insec=False
for l in dis:
    if l.startswith('//--------------------- .text.'):
        insec = kern in l
    if not insec: continue
    m=re.match(r'^(\$?_Z[^ ]*):$',l.strip())
    if m: fn=m.group(1); continue
    m=re.search(r'//## File "([^"]*)", line (\d+)',l)
    if m: fl=m.group(1).split('/')[-1]; ln=int(m.group(2)); continue
    m=re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);',l)
    if m: ins.append((fn,fl,ln,m.group(1).strip()))
print('ncu instrs',len(data),'nvdisasm instrs',len(ins))
tot=sum(int(r[ci['# Samples']]) for r in data)
byline=collections.Counter(); stall=collections.defaultdict(collections.Counter); sassof=collections.defaultdict(collections.Counter)
for (fn,fl,ln,sass),r in zip(ins,data):
    s=int(r[ci['# Samples']])
    byline[(fl,ln)]+=s
    sassof[(fl,ln)][sass[:50]]+=s
    for n in hdr:
        if n.startswith('stall_') and 'Not Issued' not in n:
            v=int(r[ci[n]])
            if v: stall[(fl,ln)][n]+=v
agg=collections.Counter()
for k in stall:
    for n,v in stall[k].items(): agg[n]+=v
print('stalls:',[(n,round(v/tot,3)) for n,v in agg.most_common(8)])
for k,v in byline.most_common(int(sys.argv[4]) if len(sys.argv)>4 else 30):
    print(f'{v/tot:.4f} {k[0]}:{k[1]} {stall[k].most_common(2)} | {sassof[k].most_common(1)[0][0]}')
