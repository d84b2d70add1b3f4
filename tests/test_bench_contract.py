"""-m "not gpu": the parts of bench.py that run without a GPU — the reference arm (oracle on the
host cores) prints exactly one JSON line with the contract's keys, rank != 0 stays silent under a
multi-rank launch, and the SURVEY §8(d) byte model behaves."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-sample-n", "6", *args], capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "newton_steps_per_s" and d["unit"] == "Newton steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] == 1 and cb["value"] == d["value"] and "Newton step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm times what it labels: the config names the SAMPLE box it really ran, with the
    # cycle it really ran (bsr.dat's V-cycle), nothing extrapolated; the bench size is only referenced
    assert d["extrapolated"] is False and d["config"]["n"] == 6 and d["config"]["amg_cycle_type"] == "V"
    assert d["config"]["dofs"] == 3 * 7 ** 3 and d["config"]["bench_workload_n"] == 256 and "6^3" in d["config"]["workload"]
    assert abs(d["us_per_dof"] - 1e3 * d["ms_per_step"] / d["config"]["dofs"]) < 1e-9 * d["us_per_dof"]
    assert len(d["seconds_per_step_each"]) == d["steps"] == 1


def test_reference_arm_shrinks_the_sample_until_k_steps_fit_the_budget():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.cpu_sample_n_for(4, 300.0, 64) == 64        # default bench.py --impl reference
    assert bench.cpu_sample_n_for(20, 300.0, 64) == 48       # the driver's --steps 20
    assert bench.cpu_sample_n_for(20, 300.0, 32) == 32
    assert bench.cpu_sample_n_for(1000, 300.0, 64) == 16


def test_reference_arm_other_ranks_exit_silently():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ["--gpus", "2"])
    assert out.returncode == 0 and out.stdout == ""


def test_byte_model_counts_the_survey_terms():
    sys.path.insert(0, ROOT)
    import bench
    st = types.SimpleNamespace(n_levels=3, level_rows=[1000, 100, 10], level_nnzb=[15000, 1700, 100])
    nv, nc = 1000, 5400
    one = bench.byte_model(st, nv, nc, 1, 30)
    two = bench.byte_model(st, nv, nc, 2, 30)
    # a second Krylov iteration adds one matvec, one cycle and the vectors of Gram-Schmidt step j = 1
    b_spmv = [15000 * 76 + 1001 * 4 + 6 * 1000 * 8, 1700 * 76 + 101 * 4 + 6 * 100 * 8]
    cycle = 3 * b_spmv[0] + 6 * (1000 + 100) * 8 + 3 * b_spmv[1] + 6 * (100 + 10) * 8
    assert two - one == b_spmv[0] + cycle + (2 * 2 + 2) * 3 * nv * 8
    # the [L | D | U] schedule reads less than the model, never more
    assert bench.byte_model(st, nv, nc, 5, 30, shortcuts=True) < bench.byte_model(st, nv, nc, 5, 30)
    # the K-cycle revisits the coarse levels
    assert bench.byte_model(st, nv, nc, 5, 30, kcycle_levels=3) > bench.byte_model(st, nv, nc, 5, 30)


def test_committed_bench_lines_carry_every_contract_key():
    """the evidence under profiles/ (written by bench.py on the B200 box) has the keys the measurement contract names:
    the base line, e2e with its copy sizes, the launch count, clocks, roofline (with a non-null traffic figure at the
    bench size) and the CPU baseline; the N > 1 lines add the parity gate and the halo-exchange path counts"""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def last_line(name):
        with open(os.path.join(root, "profiles", name)) as fh:
            return json.loads(fh.read().strip().splitlines()[-1])
    d = last_line("r02_bench_n256.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "newton_steps_per_s" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["n"] == 256 and d["config"]["dofs"] == 3 * 257 ** 3
    assert abs(d["value"] - d["steps"] / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-9 * d["value"] and d["warmup"] >= 3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * d["config"]["dofs"] == e["d2h_bytes_per_step"] and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 1000 * d["steps"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / r["ms_per_launch"] / 1e6) < 1e-6 * r["achieved"]
    assert r["traffic"] is not None and 0.8 < r["traffic"] / r["block_format"]["kernel_bytes_per_launch"] < 1.5
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] == 1 and c["value"] > 0 and "not extrapolated" in c["sample"]
    t = d["cpu_same_n"]["table"]
    assert [row["n"] for row in t] == [32, 64] and all(row["cpu_s_per_step"] > 100 * row["gpu_s_per_step_V"] for row in t)
    for name, n_gpus in (("r02_bench_4gpu_n128.json", 4), ("r02_bench_2gpu_strong_n256.json", 2)):
        m = last_line(name)
        assert m["n_gpus"] == n_gpus and m["multi_gpu_parity"]["ok"] is True
        assert m["halo_exchanges"]["peer_memory_kernel"] > 0 and m["halo_exchanges"]["nccl_send_recv"] == 0
        assert all(k > 0 for k in m["krylov_its_per_step"])
