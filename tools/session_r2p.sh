set -u
OUT=gpurun_out/r2p; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
python - <<'PY'
import ctypes, torch
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int(0)
for name, attr in (("maxPersistingL2", 108), ("l2CacheSize", 38), ("maxAccessPolicyWindow", 109)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print("==", name, v.value, "%.1f MB" % (v.value / 2**20))
PY
for wl in cfg2 h320 cfg1; do
  for v in "0 0" "0 200" "0 1000" "32 200" "64 500" "128 1000"; do
    set -- $v
    TFGNN_B200_SLEEP_CRIT=$1 TFGNN_B200_SLEEP_LONG=$2 timeout 200 $B --workload $wl --steps 15 --warmup 4 > $OUT/sleep_${wl}_$1_$2.json 2> $OUT/sleep_${wl}_$1_$2.err
    echo "== $wl sleep crit=$1 long=$2: $(grep -o '"ms_per_step": [0-9.]*' $OUT/sleep_${wl}_$1_$2.json | head -1)"
  done
done
for wl in h320 cfg1; do
  for v in "0 0" "0 1000"; do
    set -- $v
    TFGNN_B200_LIB=$PWD/gpurun_variants/libq4all.so TFGNN_B200_SLEEP_CRIT=$1 TFGNN_B200_SLEEP_LONG=$2 timeout 200 $B --workload $wl --steps 15 --warmup 4 > $OUT/q4all_${wl}_$1_$2.json 2> $OUT/q4all_${wl}_$1_$2.err
    echo "== q4all $wl sleep crit=$1 long=$2: $(grep -o '"ms_per_step": [0-9.]*' $OUT/q4all_${wl}_$1_$2.json | head -1)"
  done
done
TFGNN_B200_FUSED_TRACE=$OUT/trace_cfg2.bin timeout 200 $B --workload cfg2 --steps 1 --warmup 1 > $OUT/trace_cfg2.json 2> $OUT/trace_cfg2.err
TFGNN_B200_FUSED_TRACE=$OUT/trace_cfg2_s5.bin timeout 200 $B --workload cfg2 --steps 5 --warmup 3 > $OUT/trace_cfg2_s5.json 2> $OUT/trace_cfg2_s5.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__inst_executed.sum --clock-control none -k regex:fused_rgcn -s 3 -c 1 --csv --log-file $OUT/ncu_h320_dram.csv $B --workload h320 --steps 3 --warmup 3 > $OUT/ncu_h320_dram.log 2>&1
tail -3 $OUT/ncu_h320_dram.csv | cut -c1-300
echo "session r2p done"
