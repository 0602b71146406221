for W in 8192 16384 32768; do
for cfg in "1 0" "2 0" "2 1" "4 0"; do
  set -- $cfg
  B2GPU_GROUPS=$1 B2GPU_PIPELINE=$2 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 10 --preroll 200 --worlds-per-gpu $W 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('worlds $W groups $1 pipeline $2', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done; done
