for mb in 2 4 8 16 32 64; do VQB200_HOST_CHUNK_MB=$mb timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_mb $mb e2e ms', round(j['e2e']['ms_per_step'],3))"; done
