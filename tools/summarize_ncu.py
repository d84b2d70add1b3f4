"""Turn the scratch ncu outputs of tools/r02_evidence.sh (gpurun_out/) into the small tracked files under profiles/:
traffic JSON of the dominant sweep, per-kernel share table of the launch-list window, selected metrics of the
--set full captures.   python tools/summarize_ncu.py [round_tag]"""
import csv, gzip, json, os, sys, collections, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def ncu_rows(path):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if l.startswith('"')))
    return rows[0], rows[1:]


def short(k):
    k = k.replace("void ", "").replace("pnpb200::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    return k.split("(")[0]


# ---- dram traffic of the dominant sweep
tp = os.path.join(G, "traffic_gsd_n256.csv")
if os.path.exists(tp):
    hdr, rows = ncu_rows(tp)
    ki, mi, vi, ii, ui = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per = collections.OrderedDict()
    for r in rows:
        per.setdefault(r[ii], {"kernel": short(r[ki])})[r[mi]] = num(r[vi])
    launches = list(per.values())
    ncol = 4
    sweep = launches[-ncol:]                       # the timed sweep (the first ncol launches are the probe's warm-up sweep)
    rd = sum(l["dram__bytes_read.sum"] for l in sweep); wr = sum(l["dram__bytes_write.sum"] for l in sweep)
    us = sum(l["gpu__time_duration.sum"] for l in sweep) / 1e3
    nv = 257 ** 3; nnzb = 253036801
    alg = nnzb * 76.0 + (nv + 1) * 4.0 + 3.0 * 3 * nv * 8 + 72.0 * nv + 3 * nv * 8
    fmt = int(os.environ.get("VALUES_PER_BLOCK", "7"))
    out = {"n": 256, "values_per_block": fmt, "kernel_bytes": alg - (16.0 * nnzb if fmt == 7 else 0.0), "source": "gpurun_out/traffic_gsd_n256.csv (copied to profiles/%s_traffic_gsd_n256_launches.csv): ncu --profile-from-start off --metrics "
           "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:gs_stream_kernel -c 8, "
           "python tools/probe_rows.py 256 9 1 (tools/ncu_traffic.sh); the sweep = the last 4 launches" % tag,
           "gs_sweep": {"kernel": "gs_stream_kernel<S_GSD=5>: backward sweep + delta, %d colour launches" % ncol,
                        "dram_bytes_read": rd, "dram_bytes_written": wr, "dram_bytes_per_sweep": rd + wr, "algorithmic_bytes": alg,
                        "ratio": (rd + wr) / alg, "us_under_ncu": us, "dram_gbs_under_ncu": (rd + wr) / us / 1e3,
                        "per_launch": [{"dram_bytes": l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"], "us": l["gpu__time_duration.sum"] / 1e3} for l in sweep]}}
    json.dump(out, open(os.path.join(P, "%s_traffic_gsd_n256.json" % tag), "w"), indent=1)
    shutil.copy(tp, os.path.join(P, "%s_traffic_gsd_n256_launches.csv" % tag))
    print("traffic: %.3f GB per sweep = %.3fx algorithmic, %.1f GB/s under ncu" % ((rd + wr) / 1e9, (rd + wr) / alg, (rd + wr) / us / 1e3))

# ---- launch list window(s): per-kernel totals
for part in ("A", "B"):
    lp = os.path.join(G, "%s_launches_n256_%s.csv" % (tag, part))
    if not os.path.exists(lp) or os.path.getsize(lp) < 1000:
        continue
    hdr, rows = ncu_rows(lp)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows:
        t = tot[short(r[ki])]; t[0] += num(r[vi]) / 1e6; t[1] += 1
    allms = sum(v[0] for v in tot.values())
    with open(os.path.join(P, "%s_launches_n256_%s_summary.txt" % (tag, part)), "w") as f:
        f.write("ncu launch list, window %s of the timed step at 256^3 (%d launches, %.1f ms of gpu__time_duration; cold-cache, serialised)\n" % (part, len(rows), allms))
        f.write("%-60s %9s %7s %6s\n" % ("kernel", "ms", "calls", "share"))
        for k, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
            f.write("%-60s %9.3f %7d %5.1f%%\n" % (k[:60], ms, c, 100 * ms / allms))
    with open(lp, "rb") as fi, gzip.open(os.path.join(P, "%s_launches_n256_%s.csv.gz" % (tag, part)), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("launch list %s: %d launches, %.1f ms" % (part, len(rows), allms))

# ---- --set full captures: a few dozen metrics per launch
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
for name in ("gs_stream_n128", "asm_setup_n128"):
    rp = os.path.join(G, "%s_%s_raw.csv" % (tag, name))
    if not os.path.exists(rp) or os.path.getsize(rp) < 1000:
        continue
    with open(rp) as f:
        rows = list(csv.reader(f))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [hdr.index(c) for c in ("ID", "Kernel Name", "Block Size", "Grid Size")] + [hdr.index(k) for k in KEEP if k in hdr]
    with open(os.path.join(P, "%s_%s_metrics.csv" % (tag, name)), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols]); w.writerow([units[c] for c in cols])
        for r in body:
            w.writerow([short(r[c]) if hdr[c] == "Kernel Name" else r[c] for c in cols])
    print("%s: %d launches, %d metrics kept" % (name, len(body), len(cols) - 4))
for f in ("bench_n256.json", "kineto_step_n256.txt", "krylov_detail_n256.log", "setup_phases_n256.log"):
    src = os.path.join(G, "%s_%s" % (tag, f))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_%s" % (tag, f)))
