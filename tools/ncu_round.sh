# ncu evidence for profiles/ (one GPU, run through gpurun): launch list of the timed step at 128^3
# and a --set full capture of the fine-level row kernels
set -x
python bench.py --profile --n 128 > gpurun_out/profile_plain.json 2> gpurun_out/profile_plain.err || exit 1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv \
    --log-file gpurun_out/launches_ldu_n128.csv python bench.py --profile --n 128 > gpurun_out/ncu_launch.log 2>&1
python tools/probe_rows.py 128 9,8,10,0,7 3 > gpurun_out/probe_rows_plain.log 2>&1 || exit 1
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:bsrt_row -c 46 \
    -o gpurun_out/rowkernels_ldu_n128 python tools/probe_rows.py 128 9,8,10,0,7 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
