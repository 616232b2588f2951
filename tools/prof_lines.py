"""Per-source-line instruction counts of one profiled launch (file-aware).

usage: python tools/prof_lines.py REPORT.ncu-rep [launch_index] [top_n]
Reads the ncu source page (cuda,sass view) and prints, for OUR kernel file, warp-instructions
executed per warp-step (instructions / number of warps that run one tile) per source line, plus
the totals that come from inlined CUDA header files (shuffle/atomic intrinsics).
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
# instructions are reported per 32 agent-steps of the C3 bench shape (65536 envs x 16 agents): one warp-step
# of the lane-per-agent kernel, 1/16 of a warp's work in the env-per-thread kernel
WARPS = 32768.0

out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections = []  # (file, rows)
cur = None
for r in rows:
    if r and r[0] == 'File Path':
        cur = [r[1], []]
        sections.append(cur)
    elif cur is not None:
        cur[1].append(r)
files = []
for f, _ in sections:
    if f in files:
        break
    files.append(f)
nf = len(files)
sections = sections[launch * nf:(launch + 1) * nf]
srcs = {}
total = 0
for f, rs in sections:
    hdr = next(r for r in rs if r and r[0] == 'Line No')
    ie = hdr.index('Instructions Executed')
    isamp = hdr.index('# Samples')
    per = {}
    for r in rs:
        if r and r[0].isdigit():
            try:
                per[int(r[0])] = (int(r[ie]), int(r[isamp]))
            except ValueError:
                pass
    t = sum(v[0] for v in per.values())
    total += t
    print('== %s: %.0f i/w' % (f.split('/')[-1], t / WARPS))
    if '/csrc/' in f:
        src = srcs.setdefault(f, open('dl_reference_models_b200/csrc/' + f.split('/')[-1]).read().split('\n'))
        for ln, (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
            print('%5d %7.1f i/w %5d smp | %s' % (ln, n / WARPS, s, src[ln - 1].strip()[:110]))
print('total %.0f i/w' % (total / WARPS))
