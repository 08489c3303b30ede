mkdir -p gpurun_out
run() { echo "$@"; env "$@" timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --no-clock-sampler $WL 2>&1 | grep -o "\"ms_per_step\": [0-9.]*\|rror.*" | head -3; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused" 2>&1 | tail -5
run TFGNN_B200_FUSED_PAIR=1
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_FUSED_STAGES=4
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_FUSED_STAGES=4 TFGNN_B200_GATHER_Q=2
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_FUSED_STAGES=4 TFGNN_B200_GATHER_Q=5
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_DEBUG_SKIP=18
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_DEBUG_SKIP=22
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_DEBUG_SKIP=4
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_L2_PERSIST_MB=0
run TFGNN_B200_FUSED_PAIR=1 TFGNN_B200_RING_DISCARD=0
