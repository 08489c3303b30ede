mkdir -p gpurun_out
run() { echo "$@"; env "$@" timeout 200 python bench.py --skip-e2e --skip-cpu-baseline --no-clock-sampler $WL 2>&1 | grep -o "\"ms_per_step\": [0-9.]*\|rror.*" | head -3; }
run TFGNN_B200_FUSED_BK=32 TFGNN_B200_RING_SLOTS=4
run TFGNN_B200_FUSED_BK=32 TFGNN_B200_RING_SLOTS=5
run TFGNN_B200_FUSED_BK=32 TFGNN_B200_RING_SLOTS=4 TFGNN_B200_GATHER_Q=2 TFGNN_B200_FUSED_STAGES=3
run TFGNN_B200_FUSED_BK=32 TFGNN_B200_RING_SLOTS=4 TFGNN_B200_GATHER_Q=3
run TFGNN_B200_FUSED_BK=32 TFGNN_B200_RING_SLOTS=4 TFGNN_B200_DEBUG_SKIP=1
WL="--workload h320"
run TFGNN_B200_FUSED_BK=32
WL="--workload cfg4"
run TFGNN_B200_FUSED_BK=32
run TFGNN_B200_FUSED_BK=16
