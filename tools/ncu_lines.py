"""Per source line: instructions executed and stall samples of one kernel, from an .ncu-rep captured with --import-source on and
the line table of the object the kernel came from (cuobjdump -xelf all build/X.o; nvdisasm --print-line-info -c X.sm_100a.cubin > lines.txt).
usage: ncu_lines.py REPORT.ncu-rep MANGLED_KERNEL_NAME LINES.txt  (same build as the capture: the SASS offsets must agree)"""
import csv, re, sys, collections
rep, func, linefile = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '/tmp/exp/inflate_lines.txt')
import subprocess
import tempfile, os
tmp = os.path.join(tempfile.mkdtemp(), "src.csv")
subprocess.run(f"ncu -i {rep} --page source --csv > {tmp} 2>/dev/null", shell=True)
rows=list(csv.reader(open(tmp)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=rows[hi+1:]
ci={h:i for i,h in enumerate(hdr)}
# disassembly with line info: sequence of instructions with current file:line
lines=open(linefile).read().split('\n')
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+func+':')][0]
seq=[]; cur=('?',0)
for l in lines[start+1:]:
    if l.startswith('//--------------------- .text.'): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((int(m.group(1),16), m.group(2), cur))
print("sass instrs in disasm", len(seq), "in ncu", len(data))
base=int(data[0][0],16)
off2line={o:c for o,_,c in seq}
agg=collections.Counter(); samp=collections.Counter()
tot=0; tots=0
for r in data:
    off=int(r[0],16)-base
    c=off2line.get(off, ('?',0))
    n=int(r[ci['Instructions Executed']] or 0); s=int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
    agg[c]+=n; samp[c]+=s; tot+=n; tots+=s
print("total inst", tot, "samples", tots)
# per file region
def region(c):
    f,l=c
    return c
top=sorted(agg.items(), key=lambda kv:-kv[1])[:70]
src={}
for f in ('inflate_spec.h','inflate_spec.inc','inflate_core.h','inflate_group.inc'):
    src[f]=open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'zsc_b200', 'csrc', 'cuda', f)).read().split('\n')
for (f,l),n in top:
    text=src[f][l-1].strip()[:90] if f in src and 0<l<=len(src[f]) else ''
    print(f"{100*n/tot:5.1f}% inst {100*samp[(f,l)]/max(tots,1):5.1f}% samp  {f}:{l}  {text}")
# by file
byf=collections.Counter(); bys=collections.Counter()
for (f,l),n in agg.items(): byf[f]+=n; bys[f]+=samp[(f,l)]
print({f: (round(100*n/tot,1), round(100*bys[f]/max(tots,1),1)) for f,n in byf.items()})
# ranges in inflate_spec.h
def rng(f,a,b): return sum(n for (ff,l),n in agg.items() if ff==f and a<=l<=b)*100/tot
