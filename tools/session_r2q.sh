set -u
OUT=gpurun_out/r2q; mkdir -p $OUT
B="python bench.py --skip-e2e --skip-cpu-baseline --skip-secondary --no-clock-sampler"
for v in "1 0" "2 0" "2 6" "2 3" "1 0" "2 0"; do
  set -- $v
  TFGNN_B200_FUSED_SPLIT=$1 TFGNN_B200_RING_SLOTS=$2 timeout 200 $B --workload h320 --steps 15 --warmup 4 > $OUT/split_h320_$1_$2.json 2> $OUT/split_h320_$1_$2.err
  echo "== h320 split=$1 slots=$2: $(grep -o '"ms_per_step": [0-9.]*' $OUT/split_h320_$1_$2.json | head -1) $(tail -1 $OUT/split_h320_$1_$2.err | cut -c1-150)"
done
TFGNN_B200_FUSED_SPLIT=2 timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k "h320 or cfg5" > $OUT/pytest_scale_split.log 2>&1; echo "== scale split=2: $(tail -1 $OUT/pytest_scale_split.log)"; grep "rel err" $OUT/pytest_scale_split.log
TFGNN_B200_FUSED_SPLIT=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__cycles_active.avg,sm__inst_executed.sum,lts__t_bytes.sum --clock-control none -k regex:fused_rgcn -s 3 -c 1 --csv --log-file $OUT/ncu_h320_split.csv $B --workload h320 --steps 3 --warmup 3 > $OUT/ncu_h320_split.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open("gpurun_out/r2q/ncu_h320_split.csv")):
    if len(r) > 12 and r[0] == "0": print("==", r[4][:40], r[-3], r[-2], r[-1])
PY
TFGNN_B200_FUSED_SPLIT=2 TFGNN_B200_FUSED_TRACE=$OUT/trace_h320_split.bin timeout 200 $B --workload h320 --steps 1 --warmup 1 > $OUT/trace.json 2> $OUT/trace.err
echo "session r2q done"
