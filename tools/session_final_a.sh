# Final check of a build on one B200: every GPU suite in its own process, smoke(), the official bench line, the reference arm.
set -u
OUT=gpurun_out/${1:-final_a}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
for f in test_gpu_parity test_gpu_batch_builder test_gpu_graph_ops test_gpu_scale test_tf_golden; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q --maxfail=8 -s > $OUT/pytest_$f.log 2>&1
  echo "== $f: rc=$? $(tail -1 $OUT/pytest_$f.log)"; grep -E "^(FAILED|ERROR)|rel err" $OUT/pytest_$f.log | head -12
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "== bench rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench_full.json").read().strip().splitlines()[-1])
print("cfg2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "e2e val", d["e2e"]["value"], "clocks", d["clocks"])
for k, v in (d.get("secondary") or {}).items():
    print(k, "ms", v["ms_per_layer"], "frac", v["roofline"]["frac"], "e2e ms", v.get("e2e", {}).get("ms_per_step"), "traffic", v["roofline"]["traffic"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 600 $OUT/bench_reference.json
echo "session final_a done"
