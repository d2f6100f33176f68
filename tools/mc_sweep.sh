out=gpurun_out/sink4.txt
: > $out
for cfg in "65536 6 16" "65536 8 16" "65536 6 32" "32768 12 16" "32768 16 16"; do
  set -- $cfg
  echo "shard=$1 depth=$2 export_ctas=$3" >> $out
  MAS_B200_EXPORT_CTAS=$3 taskset -c 0-3 timeout 300 python bench.py --scaling strong --shard $1 --depth $2 --priorities 0 --steps 20 --warmup 5 --no-configs --no-parity >> $out 2>&1
done
