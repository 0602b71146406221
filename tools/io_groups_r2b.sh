for g in 4 4 5 8; do
B2GPU_IO_GROUPS=$g timeout 150 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 40 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); e=d['e2e']; print('io groups $g e2e ms/step %.3f device same regime %.3f'%(e['ms_per_step'], e['device_resident_ms_per_step_same_regime']))"
done
