mkdir -p gpurun_out/sens
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "fused" 2>&1 | grep -E "^E  +|Error|FAILED|passed|failed|py:[0-9]+" | head -12
VARIANTS=10 SENS_SLABS=0 B=4 python scripts/tune_trilinear.py > gpurun_out/sens/tri_b4.log 2>&1; tail -3 gpurun_out/sens/tri_b4.log
VARIANTS=26 SENS_SLABS=0,8,16,32 B=16 python scripts/tune_trilinear.py > gpurun_out/sens/tri_b16.log 2>&1; tail -6 gpurun_out/sens/tri_b16.log
