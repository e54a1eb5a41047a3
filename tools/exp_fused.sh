for cfg in "16 2" "16 3" "8 2"; do
  set -- $cfg
  FH_PRO_NB=$1 FH_PRO_STAGES=$2 timeout 300 python bench.py --steps 2 --warmup 3 --no-latency --no-cpu-baseline --breakdown gpurun_out/r2_exp_$1_$2.json > gpurun_out/r2_exp_$1_$2.out 2> gpurun_out/r2_exp_$1_$2.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_exp_$1_$2.out').read().strip().splitlines()[-1]); print('NB $1 S $2', round(d['ms_per_step'],1), d['stage_ms']['tc_conv'])"
done
