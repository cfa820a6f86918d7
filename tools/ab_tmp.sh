for v in texturefusion_b200/libtexfusion_b200.so build/variants/dyn.so; do
  echo "== lib=$v"
  for a in "" "--steps 20 --warmup 5"; do
  TEXFUSION_B200_LIB=$PWD/$v python bench.py --no-cpu-baseline $a 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f e2e %.0f launch %.1f frac %.3f hash %s' % (d['value'], d['e2e']['value'], r['avg_launch_us'], r['frac'], d['map_hash']), d['stage_us_per_frame'])"
  done
done
