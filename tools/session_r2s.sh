set -u
OUT=gpurun_out/r2s; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
run() { # name workload env...
  local name=$1 wl=$2; shift 2
  env "$@" timeout 200 $B --workload $wl --steps 15 --warmup 4 > $OUT/$name.json 2> $OUT/$name.err
  echo "== $name: $(grep -o '"ms_per_step": [0-9.]*' $OUT/$name.json | head -1) $(tail -1 $OUT/$name.err | cut -c1-160)"
}
run cfg1_base cfg1 X=1
run cfg1_bk16_q6 cfg1 TFGNN_B200_FUSED_BK=16 TFGNN_B200_GATHER_Q=6
run cfg1_bk16_q4 cfg1 TFGNN_B200_FUSED_BK=16 TFGNN_B200_GATHER_Q=4
run cfg1_bk16_q8 cfg1 TFGNN_B200_FUSED_BK=16 TFGNN_B200_GATHER_Q=8
run cfg1_nosplit cfg1 TFGNN_B200_FUSED_SPLIT=0
run cfg1_base2 cfg1 X=1
run h320_base h320 X=1
run h320_bk16 h320 TFGNN_B200_FUSED_BK=16
run h320_bk16_s4 h320 TFGNN_B200_FUSED_BK=16 TFGNN_B200_FUSED_STAGES=4
run h320_q3 h320 TFGNN_B200_GATHER_Q=3
run h320_split6 h320 TFGNN_B200_FUSED_SPLIT=2 TFGNN_B200_RING_SLOTS=6
run h320_base2 h320 X=1
run cfg2_base cfg2 X=1
run cfg2_q3 cfg2 TFGNN_B200_GATHER_Q=3
run cfg2_slots5 cfg2 TFGNN_B200_RING_SLOTS=5
run cfg2_base2 cfg2 X=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_cfg3.csv $B --workload cfg3 --steps 2 --warmup 2 > $OUT/launches_cfg3.log 2>&1
python tools/launch_summary.py $OUT/launches_cfg3.csv 2>/dev/null | tail -14
echo "session r2s done"
