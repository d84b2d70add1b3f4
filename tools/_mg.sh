run() { echo "== $*"; env "$@" PNPB200_HALO_IPC=0 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29730 tools/mg_determinism.py 96 8 2>&1 | grep "DISTINCT\|Error\|error" | tail -3; }
run PNPB200_GS_STREAM=0
run DET_SHORTCUTS=0
run PNPB200_COARSE_GRAPH=0
run DET_CYCLE=1
