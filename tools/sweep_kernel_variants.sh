#!/bin/bash
for f in scratch/lib_*.so; do
  for ctas in 0 4; do
    out=$(MAPF_B200_LIB=$PWD/$f MAPF_STEP_CTAS_PER_SM=$ctas python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms kernel, %.4f ms/step, value %.3e' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value']))")
    echo "$f ctas_per_sm=$ctas : $out"
  done
done
