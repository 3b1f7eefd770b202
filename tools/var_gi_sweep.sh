for gi in 2 4 8 16; do
  GPR_VAR_GI=$gi python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gi$gi.json 2>/dev/null
  python -c "
import json; b=json.load(open('gpurun_out/bench_gi$gi.json')); print('GI=$gi', b['value'], b['roofline']['achieved'])"
done
for gi in 8 16; do
  GPR_VAR_GI=$gi python tools/prof_target.py > /dev/null 2>&1 && GPR_VAR_GI=$gi ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:var_tiles -c 1 python tools/prof_target.py 2>&1 | grep -E "dram__bytes_read|gpu__time|hit_rate" | sed "s/^/GI=$gi /"
done
