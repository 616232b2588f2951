import csv,sys,subprocess,re
rep=sys.argv[1]; nl=int(sys.argv[2]) if len(sys.argv)>2 else 2
src=open('dl_reference_models_b200/csrc/mapf_kernels.cuh').read().split('\n')
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[2]; ie=hdr.index('Instructions Executed'); ss=hdr.index('# Samples')
per={}; samp={}
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            per[int(r[0])]=per.get(int(r[0]),0)+int(r[ie]); samp[int(r[0])]=samp.get(int(r[0]),0)+int(r[ss])
        except: pass
tot=sum(per.values()); ts=sum(samp.values())
# regions by marker comments / function starts in the source
marks=[]
for i,l in enumerate(src,1):
    m=re.match(r'\s*// -{10,} (.*)',l)
    if m: marks.append((i,m.group(1)))
    m=re.match(r'(?:template.*\n)?__device__ __forceinline__ \S+ (\w+)\(',l)
    if m: marks.append((i,'fn '+m.group(1)))
    if l.startswith('__global__') or '__global__ void' in l: marks.append((i,'kernel '+l[:60]))
    if l.startswith('struct Philox'): marks.append((i,'philox'))
marks.sort()
warps=32768
print('total instr/warp %.0f'%(tot/nl/warps))
for (lo,name),(hi,_) in zip(marks,marks[1:]+[(10**6,'')]):
    s_=sum(v for k,v in per.items() if lo<=k<hi); q=sum(v for k,v in samp.items() if lo<=k<hi)
    if s_: print('%5d %-45s %6.1f%% %7.0f i/w  samples %5.1f%%'%(lo,name[:45],100*s_/tot,s_/nl/warps,100*q/ts))
