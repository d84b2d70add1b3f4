# round-2 evidence, one GPU (every ncu command only after the plain command exited 0; outputs kept under 64 MiB)
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gputests.log 2>&1 || { tail -n 30 $O/r02_gputests.log; exit 1; }
python bench.py > $O/r02_bench_n256.json 2> $O/r02_bench_n256.err || exit 1
python tools/krylov_detail.py 256 > $O/r02_krylov_detail_n256.log 2>&1
python tools/setup_phases.py 256 > $O/r02_setup_phases_n256.log 2>&1
python tools/kineto_step.py 256 > $O/r02_kineto_step_n256.txt 2> $O/r02_kineto.err
NLAUNCH=8 bash tools/ncu_traffic.sh 256
python bench.py --profile --n 256 > $O/r02_profile_plain.json 2> $O/r02_profile_plain.err || exit 1
# launch list of the timed step in two windows (a whole step is ~9 000 launches = 15 minutes under ncu):
# A = assembly + AMG setup + the first Krylov iterations
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1900 --csv \
    --log-file $O/r02_launches_n256_A.csv python bench.py --profile --n 256 > $O/r02_ncu_launch_A.log 2>&1
echo "launch list A exit $?"
python tools/probe_rows.py 128 9,7 1 > $O/r02_probe_plain.log 2>&1 || exit 1
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gs_stream_kernel -c 8 \
    -o $O/r02_gs_stream_n128 python tools/probe_rows.py 128 9,7 1 > $O/r02_ncu_full_gs.log 2>&1
echo "full gs exit $?"
ncu -i $O/r02_gs_stream_n128.ncu-rep --page raw --csv > $O/r02_gs_stream_n128_raw.csv 2>/dev/null
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k regex:"cell_jacobian|assemble_pairs|cell_residual|residual_gather|rapw_numeric|rapw_symbolic|condense" -c 9 \
    -o $O/r02_asm_setup_n128 python bench.py --profile --n 128 > $O/r02_ncu_full_asm.log 2>&1
echo "full asm exit $?"
ncu -i $O/r02_asm_setup_n128.ncu-rep --page raw --csv > $O/r02_asm_setup_n128_raw.csv 2>/dev/null
du -sm $O; ls -la $O
# stay under the 64 MiB pull limit: the raw csv pages carry the numbers, the reports are dropped largest first
for f in $(ls -S $O/*.ncu-rep); do if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $f; fi; done
du -sm $O
