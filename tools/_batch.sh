O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gputests_c.log 2>&1; tail -n 5 $O/r02_gputests_c.log
timeout 300 python tools/probe_rows.py 256 9,7,8,10,0,1 10 > $O/probe_pack7.log 2>&1; cat $O/probe_pack7.log
PNPB200_PACK7=0 timeout 300 python tools/probe_rows.py 256 9,7,8,10,0,1 10 > $O/probe_pack9.log 2>&1; cat $O/probe_pack9.log
timeout 300 python tools/krylov_detail.py 256 > $O/kd_pack7.log 2>&1; head -14 $O/kd_pack7.log
