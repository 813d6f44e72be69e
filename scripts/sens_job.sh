mkdir -p gpurun_out/sens
VARIANTS=0 SVARIANTS=${SV:-0,1,9,13,14,15,16,17,18,19,20,21} SENS_ONLY=1 python scripts/tune_siddon.py > gpurun_out/sens/tune.log 2>&1
tail -16 gpurun_out/sens/tune.log
