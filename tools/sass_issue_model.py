"""Static in-order issue model of a SASS address range (one warp per sub-partition): fixed latencies, FP64 issue every 2 cycles.
   cuobjdump -sass -fun <kernel> lib.so > k.sass ; python tools/sass_issue_model.py k.sass 0x<lo> 0x<hi>   (used for DESIGN.md section 8)"""
import re,sys
def load(path, lo, hi):
    ins=[]
    for ln in open(path):
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',ln)
        if not m: continue
        a=int(m.group(1),16)
        if lo<=a<hi: ins.append((a,m.group(2).strip()))
    return ins
LAT={'DFMA':8,'DMUL':8,'DADD':8,'LDS':29,'LDTM':20,'MUFU':20,'IMAD':5,'MOV':5,'DSETP':12,'LDL':40,'LDG':300}
def regs(tok):
    out=[]
    for m in re.finditer(r'\bR(\d+)\b',tok):
        out.append(int(m.group(1)))
    return out
def sim(ins, wide64=True):
    ready={}; t=0; total_stall=0; chains=0
    fp_last=-10
    for a,s in ins:
        s2=re.sub(r'^@!?U?P\d+\s+','',s)
        op=s2.split()[0].split('.')[0]
        args=s2[len(s2.split()[0]):]
        parts=[p.strip() for p in args.split(',')]
        if not parts or parts==['']: t+=1; continue
        dst=regs(parts[0]); srcs=[r for p in parts[1:] for r in regs(p)]
        is64 = op in ('DFMA','DMUL','DADD','LDS','DSETP') or (op in ('LDL','LDG','STL') and '.64' in s2.split()[0])
        def expand(rs): 
            o=set()
            for r in rs:
                o.add(r); 
                if is64: o.add(r+1)
            return o
        if op in ('STS','STTM','STL','STG'): srcs=[r for p in parts for r in regs(p)]; dst=[]
        need=max([ready.get(r,0) for r in expand(srcs)]+[0])
        issue=max(t+1,need)
        if op in('DFMA','DMUL','DADD'): issue=max(issue,fp_last+2); fp_last=issue
        total_stall+=issue-(t+1)
        t=issue
        lat=LAT.get(op,4)
        for r in expand(dst): ready[r]=t+lat
    return t,total_stall
if __name__=='__main__':
    path,lo,hi=sys.argv[1],int(sys.argv[2],16),int(sys.argv[3],16)
    ins=load(path,lo,hi)
    n=len(ins)
    print('instr',n)
    for a,b in ((0,n),(0,n-450),(n-450,n)):
        t,st=sim(ins[a:b]); print(f'[{a}:{b}] cycles {t} stalls {st}')
