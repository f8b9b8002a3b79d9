# A/B of diagnostic switches on the bench scene: bash tools/ab.sh "<env assignments>" ...   (each configuration twice)
run() { env $1 python bench.py --steps ${STEPS:-60} --warmup 5 --no-cpu-baseline --stage-steps 0 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'its', d['newton_iterations'], 'ms/it', round(d['solve_gpu_ms_per_iteration'],4))"; }
for cfg in "$@"; do run "$cfg"; run "$cfg"; done
