# N = 1 headline shard, 20 steps: e2e pipelines x stagger
out=gpurun_out/e2e_n1.txt
: > $out
for cfg in "4 1" "4 0.5" "4 1.5" "5 1" "6 1" "3 1" "8 1" "4 0"; do
  set -- $cfg
  echo "e2e_depth=$1 e2e_stagger=$2" >> $out
  timeout 300 python bench.py --e2e-depth $1 --depth $(( $1 > 4 ? $1 : 4 )) --e2e-stagger $2 --steps 20 --warmup 5 --no-configs --no-parity 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(f\"   resident {d['value']/1e6:.2f} M/s  e2e {d['e2e']['value']/1e6:.2f} ms {d['e2e']['ms_per_step']:.3f} wall {d['e2e']['wall_ms_per_step']:.3f} ctrl-only {d['e2e_controls_only']['value']/1e6:.2f}\")
" >> $out
done
