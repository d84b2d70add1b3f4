O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 4 --warmup 3 > $O/r02_bench_2gpu_n256.json 2> $O/r02_bench_2gpu_n256.err
echo "exit $?"; tail -n 3 $O/r02_bench_2gpu_n256.err
