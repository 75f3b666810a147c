import re,sys,subprocess
def funcs(path):
    out=subprocess.run(["cuobjdump","-sass",path],capture_output=True,text=True).stdout
    cur=None; d={}
    for l in out.split('\n'):
        m=re.search(r'Function : (\S+)',l)
        if m: cur=m.group(1); d[cur]=[]; continue
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
        if m and cur: d[cur].append((int(m.group(1),16),m.group(2).strip()))
    return d
def loops(ins):
    res=[]
    for a,t in ins:
        if 'BRA' in t:
            m=re.search(r'0x([0-9a-f]+)',t)
            if m and int(m.group(1),16)<a:
                tgt=int(m.group(1),16)
                body=[x for x in ins if tgt<=x[0]<=a]
                res.append((tgt,a,body))
    return res
def analyze(body):
    # returns counts under models
    prev={}  # slot -> reg held in reuse cache
    n=len(body); nf=0; c1=0; c2=0; c3=0
    for a,t in body:
        t2=re.sub(r'^@!?U?P\d+\s+','',t)
        op=t2.split()[0]
        args=t2[len(op):].split(',')
        args=[x.strip() for x in args]
        srcs=args[1:]
        reads=[]; newprev={}
        for slot,s in enumerate(srcs):
            m=re.match(r'^-?\|?(R\d+)\|?(\.reuse)?$',s)
            if not m: continue
            r=int(m.group(1)[1:])
            if prev.get(slot)!=r: reads.append(r)
            if m.group(2): newprev[slot]=r
        prev=newprev
        if op.startswith('FFMA') or op.startswith('FADD') or op.startswith('FMUL'): nf+=1
        ev=len(set(r for r in reads if r%2==0)); od=len(set(r for r in reads if r%2==1))
        c1+=max(1,ev,od)
        c2+=max(1,len(set(reads))-1)
        c3+=max(1,ev,od) if op.startswith('F') else 1
    return n,nf,c1,c2
if __name__=="__main__":
    d=funcs(sys.argv[1])
    for name,ins in d.items():
        if len(sys.argv)>2 and sys.argv[2] not in name: continue
        for tgt,a,body in loops(ins):
            n,nf,c1,c2=analyze(body)
            if nf>20: print(f"{name[:40]:40s} loop {tgt:05x}-{a:05x} n={n} nf={nf} bankmodel={nf/c1:.3f} Hmodel={nf/c2:.3f}")

def analyze2(body):
    prev={}; tot=0; same3=0; nf=0; r3=0; hist={}
    for a,t in body:
        t2=re.sub(r'^@!?U?P\d+\s+','',t)
        op=t2.split()[0]
        args=[x.strip() for x in t2[len(op):].split(',')]
        srcs=args[1:]
        reads=[]; newprev={}
        for slot,s in enumerate(srcs):
            m=re.match(r'^-?\|?(R\d+)\|?(\.reuse)?$',s)
            if not m: continue
            r=int(m.group(1)[1:])
            if prev.get(slot)!=r: reads.append(r)
            if m.group(2): newprev[slot]=r
        prev=newprev
        if not op.startswith('FFMA'): continue
        nf+=1
        rs=set(reads); tot+=len(rs)
        ev=len([r for r in rs if r%2==0]); od=len(rs)-ev
        key=(ev,od); hist[key]=hist.get(key,0)+1
    return nf,(tot/nf if nf else 0),hist
