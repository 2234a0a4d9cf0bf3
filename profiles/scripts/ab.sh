#!/bin/bash
# A/B of builds of the library on the headline step: bash profiles/scripts/ab.sh [size] [lib ...]
# (default: the shipped library against pyflwdir_b200/libpfd_b200_exp.so)
SIZE=${1:-8192}
shift
LIBS=("$@")
[ ${#LIBS[@]} -eq 0 ] && LIBS=(pyflwdir_b200/libpfd_b200_exp.so)
for lib in "" "${LIBS[@]}"; do
  PFD_B200_LIB=$lib python bench.py --size $SIZE --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --extras none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${lib:-baseline}', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
done
