O=gpurun_out; mkdir -p $O; T=r01j
timeout 600 python -m pytest tests -m gpu -q --timeout 180 -k "evaluate or golden or pipeline or model" > $O/${T}_pytest_eval.txt 2>&1; tail -3 $O/${T}_pytest_eval.txt | cut -c1-300
run() { timeout 300 python bench.py --no-cpu-baseline "$@" > $O/${T}_b.json 2> $O/${T}_b.err; python -c "
import json; d=json.load(open('$O/${T}_b.json')); print('$*', '| train %.1f e2e %.1f infer %.1f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['inference']['ms_per_step']), d['clocks'])"; }
run --sampler nvml
run --sampler smi
run --sampler off
run --sampler nvml --switch-interval 0.0002
run --sampler nvml --streams 3
run --sampler nvml --streams 6
run --sampler nvml --steps 10
