for v in "UVS_PROJ_OCC=5" "UVS_LINE_OCC=4" "UVS_PROJ_OCC=5 UVS_LINE_OCC=4"; do
  env $v python bench.py --no-cpu --steps 12 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value']), round(r['frac'],4), r['ms_per_launch'], {k:(v['ms_alone'],v['frac']) for k,v in r['per_kernel'].items()})"
done
