mkdir -p gpurun_out/sens
VARIANTS=0 SVARIANTS=${SV:-0,23,24,25,26,27,28,29,30,31,0} SENS_ONLY=1 python scripts/tune_siddon.py > gpurun_out/sens/tune2.log 2>&1
tail -13 gpurun_out/sens/tune2.log
python scripts/registration_bench.py --steps 1000 --graph > gpurun_out/sens/reg_graph.log 2>&1; tail -1 gpurun_out/sens/reg_graph.log | cut -c1-220
python scripts/registration_bench.py --steps 300 > gpurun_out/sens/reg_eager.log 2>&1; tail -1 gpurun_out/sens/reg_eager.log | cut -c1-220
python scripts/registration_bench.py --steps 500 --graph --renderer trilinear > gpurun_out/sens/reg_graph_tri.log 2>&1; tail -1 gpurun_out/sens/reg_graph_tri.log | cut -c1-220
