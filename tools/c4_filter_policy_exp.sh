for cfg in "12 0.005" "48 0.05" "64 0.1"; do set -- $cfg
  SOM_B200_FILTER_MAXC=$1 SOM_B200_FILTER_MAXO=$2 timeout 300 python bench.py --workload c4 --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=j['roofline']
print('maxc $1 maxo $2: ms/step %.2f kernel_ms %.2f'%(j['ms_per_step'], r['kernel_ms']), j.get('filter_path',{}).get('epochs_on_it_so_far'), j.get('filter_path',{}).get('last_full_run_overflow_frac_and_candidates_per_row'), j.get('filter_path',{}).get('last_probe_overflow_frac_and_candidates_per_row'))"
done
