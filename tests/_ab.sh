set -x
timeout 600 python -m pytest tests/test_gpu_network.py -m gpu -x -q > gpurun_out/r3b_pytest_network.log 2>&1; tail -3 gpurun_out/r3b_pytest_network.log
for i in 1 2; do
B200_TIME=30 timeout 120 python tests/gpu_onestep.py bf16 65536 8 2>&1 | tail -1
B200_TIME=30 B200_WGRAD_SLICE=64 timeout 120 python tests/gpu_onestep.py bf16 65536 8 2>&1 | tail -1
B200_TIME=30 B200_WGRAD_SLICE=16 timeout 120 python tests/gpu_onestep.py bf16 65536 8 2>&1 | tail -1
B200_TIME=30 BILINEAR_B200_LIB=$PWD/bilinear_b200/lib/libbilinear_b200_base.so timeout 120 python tests/gpu_onestep.py bf16 65536 8 2>&1 | tail -1
done
