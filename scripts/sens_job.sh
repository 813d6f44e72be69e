mkdir -p gpurun_out/sens
python scripts/registration_bench.py --steps 300 > gpurun_out/sens/reg_eager.log 2>&1; tail -2 gpurun_out/sens/reg_eager.log
python scripts/registration_bench.py --steps 1000 --graph > gpurun_out/sens/reg_graph.log 2>&1; tail -2 gpurun_out/sens/reg_graph.log
python scripts/registration_bench.py --steps 300 --graph --renderer trilinear > gpurun_out/sens/reg_graph_tri.log 2>&1; tail -2 gpurun_out/sens/reg_graph_tri.log
VARIANTS=26 SENS_SLABS=0 B=64 FWD_ONLY=1 python scripts/tune_trilinear.py > gpurun_out/sens/tri_b64.log 2>&1; tail -3 gpurun_out/sens/tri_b64.log
