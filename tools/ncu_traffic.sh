# dram traffic of the dominant kernel (backward Gauss-Seidel sweep + delta: gs_stream_kernel<S_GSD>, one launch per
# colour) at the bench size: one ncu pass with three metrics (no --set full: a single pass needs no replay)
N=${1:-256}
python tools/probe_rows.py $N 9 1 > gpurun_out/traffic_plain_$N.log 2>&1 || exit 1
timeout 420 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --cache-control none -k regex:"gs_stream_kernel|bsrt_row_kernel" -c ${NLAUNCH:-8} --csv --log-file gpurun_out/traffic_gsd_n$N.csv \
    python tools/probe_rows.py $N 9 1 > gpurun_out/traffic_ncu_$N.log 2>&1
echo "ncu exit $?"
tail -n 2 gpurun_out/traffic_plain_$N.log; tail -n 2 gpurun_out/traffic_ncu_$N.log
wc -l gpurun_out/traffic_gsd_n$N.csv
