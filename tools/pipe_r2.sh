for cfg in "2 0" "2 1" "4 0" "4 1" "3 0" "8 0"; do set -- $cfg
B2GPU_GROUPS=$1 B2GPU_PIPELINE=$2 timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('groups $1 pipeline $2 ms/step %.3f'%d['ms_per_step'])"
done
