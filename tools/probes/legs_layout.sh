for lay in interleaved planar interleaved planar; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-pipeline --no-sustained --legs-layout $lay 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); L=d['legs']
print('$lay', 'global_minmax ms %.3f frac %.3f'%(L['global_minmax']['ms_per_step'],L['global_minmax']['frac_of_hbm_peak_this_rank']), 'config4 with_gen %.3f s pipeline_only %.4f s kept %d'%(L['config4_100k']['with_generation']['seconds'],L['config4_100k']['pipeline_only']['seconds'],L['config4_100k']['with_generation']['kept_windows']), L['config4_100k']['global_minmax'])
"
done
