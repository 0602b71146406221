for args in "--toi" "--toi --sleep" "--toi --workload mixed"; do
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 0 $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); k=d['roofline']['kernels_ms']; print('$args', 'ms/step %.3f'%d['ms_per_step'], ' '.join(f\"{a.replace('k_','')}={b:.3f}\" for a,b in k.items()))"
done
