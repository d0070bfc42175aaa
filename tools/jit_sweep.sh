for cfg in "32 8 2" "16 8 3" "16 4 3" "8 8 3" "8 4 4" "16 8 2"; do set -- $cfg; PFFRG_JIT_CHUNK=$1 PFFRG_JIT_ACC=$2 PFFRG_JIT_MINBLOCKS=$3 timeout 200 python bench.py --steps 3 --warmup 3 --synthetic-state --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$cfg', round(d['breakdown_ms']['ms_v4_flow'],2))"; done
