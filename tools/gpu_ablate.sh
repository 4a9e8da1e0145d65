#!/bin/bash
# time the bench with alternative builds of the library (build/abl/*.so): where does the fused kernel's time go?
for f in build/abl/*.so; do
  CHALTE_LIB=$PWD/$f python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$f', 'step_ms', round(d['ms_per_step'],4), 'fused_ms', round(d['roofline']['avg_launch_ms'],4))"
done
