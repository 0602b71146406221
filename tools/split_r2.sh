for t in 256 296 216 320 192 148; do
B2GPU_GROUP0_TILES=$t timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('group0 tiles $t ms/step %.3f'%d['ms_per_step'])"
done
