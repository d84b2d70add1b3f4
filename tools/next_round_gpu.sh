# First GPU call of the next round (one GPU, ~4 box-minutes): evidence for the state the last round ended in.
#  1. bench line of the default configuration (4-colour BoxMesh sweeps, coarse-level CUDA graph)
#  2. the same with the graph colouring / without the graph, for the A/B table in profiles/README.md
#  3. single-pass ncu capture of the dram traffic of one backward sweep (4 colour launches) at 256^3
set -x
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
PNPB200_BOX_COLORING=0 python bench.py --steps 3 > gpurun_out/bench_jp_colouring.json 2> gpurun_out/bench_jp_colouring.err
PNPB200_COARSE_GRAPH=0 python bench.py --steps 3 > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err
NLAUNCH=8 bash tools/ncu_traffic.sh 256
