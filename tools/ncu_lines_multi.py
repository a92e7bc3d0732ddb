"""Per source line of EVERY kernel in a report: instructions executed and stall samples (see ncu_lines.py for the inputs).
usage: ncu_lines_multi.py REPORT.ncu-rep LINES.txt SOURCE_DIR [TOP]   (LINES.txt = nvdisasm --print-line-info -c of the same build)"""
import csv, re, sys, collections, subprocess, tempfile, os
rep, linefile, srcdir = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = os.path.join(tempfile.mkdtemp(), "src.csv")
subprocess.run(f"ncu -i {rep} --page source --csv > {tmp} 2>/dev/null", shell=True)
rows = list(csv.reader(open(tmp)))
# line table: per .text section, list of (offset, (file, line))
secs = {}; cur_sec = None; cur = ('?', 0)
for l in open(linefile):
    m = re.match(r'\.text\.(\S+):', l)
    if m: cur_sec = m.group(1); secs[cur_sec] = {}; cur = ('?', 0); continue
    if cur_sec is None: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: secs[cur_sec][int(m.group(1), 16)] = cur
srcs = {}
def text(f, l):
    if f not in srcs:
        p = os.path.join(srcdir, f)
        srcs[f] = open(p).read().split('\n') if os.path.exists(p) else []
    return srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ''
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; ci = {h: k for k, h in enumerate(hdr)}
        j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'): data.append(rows[j]); j += 1
        data = [r for r in data if r and re.match(r'^(0x)?[0-9a-f]+$', r[0])]
        # section with the same instruction count
        cand = [s for s, t in secs.items() if len(t) == len(data)]
        short = re.sub(r'\(.*', '', name).replace('void ', '')
        base_name = re.sub(r'<.*', '', short)
        cand = [s for s in cand if base_name in s] or cand
        table = secs[cand[0]] if cand else {}
        base = int(data[0][0], 16)
        agg = collections.Counter(); samp = collections.Counter(); tot = tots = 0
        for r in data:
            c = table.get(int(r[0], 16) - base, ('?', 0))
            n = int(r[ci['Instructions Executed']] or 0); s = int(r[ci['Warp Stall Sampling (All Samples)']] or 0)
            agg[c] += n; samp[c] += s; tot += n; tots += s
        print(f"==== {short}: {tot} warp instructions, {tots} samples ({cand[0] if cand else 'no line table'})")
        keys = sorted(set(agg) | set(samp), key=lambda c: -(samp[c]))[:top_n]
        for c in keys:
            print(f"{100*agg[c]/max(tot,1):5.1f}% inst {100*samp[c]/max(tots,1):5.1f}% samp  {c[0]}:{c[1]}  {text(*c)}")
        i = j
    else: i += 1
