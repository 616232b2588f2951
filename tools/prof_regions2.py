"""Stall samples and instructions per region of mapf_env_kernel.cuh from an ncu report (regions found by marker comments).
usage: python tools/prof_regions2.py REPORT.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'File Path':
        cur = [r[1], []]; secs.append(cur)
    elif cur is not None:
        cur[1].append(r)
src = open('dl_reference_models_b200/csrc/mapf_env_kernel.cuh').read().split('\n')
marks = [('helpers', 'template <bool VEC>'), ('prologue', '__global__ void __launch_bounds__(448, 1) mapf_step_env_kernel'),
         ('emit_final', 'auto emit_final = '), ('tile setup', 'for (int tile = blockIdx.x'),
         ('prepass', 'pre-pass: agent records'), ('walk A', 'uint32_t moved_m = 0, failed_m = 0'),
         ('phase B', '// Phase B --'), ('flush', 'coalesced flush of the stage rows'),
         ('words/degen', 'the rest of the env words'), ('lifelong', 'lifelong goal reassignment (ENV:284'),
         ('epilogue', 'epilogue: owner masks'), ('lock result', 'lock detection result'), ('rewards', 'rewards & termination'),
         ('episode end+reset', 'episode end: metrics'), ('writeback', 'env words write-back')]
pos = []
for name, m in marks:
    ln = next(i + 1 for i, l in enumerate(src) if m in l)
    pos.append((ln, name))
pos.sort()
for f, rs in secs:
    if 'mapf_env_kernel' not in f:
        continue
    hdr = next(r for r in rs if r and r[0] == 'Line No')
    ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
    per = {}
    for r in rs:
        if r and r[0].isdigit():
            try:
                per[int(r[0])] = (int(r[ie]), int(r[isamp]))
            except ValueError:
                pass
    tot_s = sum(v[1] for v in per.values()); tot_i = sum(v[0] for v in per.values())
    print('file total: %.1f i/w, %d samples' % (tot_i / 32768.0, tot_s))
    for k, (ln, name) in enumerate(pos):
        end = pos[k + 1][0] - 1 if k + 1 < len(pos) else 10 ** 9
        i = sum(v[0] for l, v in per.items() if ln <= l <= end) / 32768.0
        s = sum(v[1] for l, v in per.items() if ln <= l <= end)
        print('%-20s lines %4d-%-5s %7.1f i/w %6d smp %5.1f%%' % (name, ln, end if end < 10 ** 8 else 'end', i, s, 100.0 * s / tot_s))
    break
