#!/bin/bash
# config 4 at epochs 20..24 of the 100-epoch schedule (15 - 22 candidates per row) with the policy threshold at 12 / 24 / 32
for maxc in 12 24 32; do
  SOM_B200_FILTER_MAXC=$maxc timeout 300 python bench.py --workload c4 --no-cpu --no-extra --warmup 20 --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=j['roofline']
print('maxc $maxc: ms/step %.2f kernel_ms %.2f'%(j['ms_per_step'], r['kernel_ms']), j.get('filter_path',{}).get('epochs_on_it_so_far'), j.get('filter_path',{}).get('last_full_run_overflow_frac_and_candidates_per_row'), j.get('filter_path',{}).get('last_probe_overflow_frac_and_candidates_per_row'))"
done
