for c in 25 50 100 200 0; do echo "CHUNK=$c"; GASR_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages_ms_per_step'], d['pipeline']['launches_per_stage'])"; done
