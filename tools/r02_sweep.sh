# BASELINE configs[4]: the size sweep on one GPU (resident + end-to-end steps/s, stage times, roofline fractions per size)
O=gpurun_out
for n in 64 96 128 192 320; do
  timeout 600 python bench.py --n $n --steps 3 --warmup 3 --cpu-sample-n 16 > $O/r02_sweep_n$n.json 2> $O/r02_sweep_n$n.err
  echo "n=$n exit $?"; tail -n 3 $O/r02_sweep_n$n.err
done
PNPB200_SHARE_SCRATCH=1 timeout 300 python bench.py --n 128 --steps 3 --warmup 3 --cpu-sample-n 16 > $O/r02_sweep_n128_shared.json 2> $O/r02_sweep_n128_shared.err
echo "shared exit $?"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
