"""-m "not gpu": the host mirrors of the reference's plain-C++ control logic on the path:
Newton_Status (src/newton_status.cpp:11-100 -> modular-pnp_b200/host/newton_status.hpp) and the domain parameter reader
/ domain_build (src/domain.cpp:18-311 -> host/domain.hpp), against
  * tests/golden/host_logic.json: outputs of the REFERENCE's own two source files compiled where they lie
    (oracle/_ref/libpnpref_host.so = oracle/Makefile + oracle/ref_wrappers_host.inc + oracle/shim/dolfin.h), written by
    `python tests/test_host_logic.py --write-golden` in the authoring container, and
  * that library directly where it is present (it travels with the repo; /root/reference itself is not read).
Status convention: where the reference ends the process through fasp_chkerr(status < 0), the wrapper reports that
status; the mirror RETURNS it (host/domain.hpp header)."""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "host_logic.json")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpnpref_host.so")
HOST = os.path.join(ROOT, "modular-pnp_b200", "host")

# --- inputs -----------------------------------------------------------------------------------------------------------
# (max_iterations, initial_residual, rel_tol, max_tol, [(l2 residual, max residual), ...]) — the loop of
# benchmarks/pnp_exact_solution/main.cpp:334-356: while needs_to_iterate: solve; update_residuals; update_iteration
NEWTON_CASES = [
    (15, 3.2e-1, 1e-8, 1e-10, [(3.2e-1 * 10.0 ** (-1.3 * k), 0.5 * 10.0 ** (-1.1 * k)) for k in range(1, 20)]),   # configs[0]: converges on rel
    (15, 1.0, 1e-8, 1e-10, [(0.9 ** k, 0.8 ** k) for k in range(1, 40)]),                                         # too slow: runs out of iterations
    (5, 2.0, 1e-3, 1e-1, [(1.0, 0.5), (0.5, 0.09), (0.1, 0.01)]),                                                # max residual accepts first
    (3, 1.0, 1e-8, 1e-10, [(1e-9, 1.0)]),                                                                        # rel accepted at once
    (1, 1.0, 1e-8, 1e-10, [(0.5, 0.5), (0.25, 0.25)]),                                                           # a single allowed iteration
    (4, 1.0, 1e-4, 1e-4, [(1e-4, 1e-4), (1e-4, 1e-4), (9e-5, 2e-4)]),                                            # exactly AT the tolerances (strict inequalities)
    (6, 5.0, 1e-8, 1e-10, [(float("nan"), 1.0), (1.0, float("nan")), (1e-12, 1e-12)]),                           # NaN residuals compare false
    (0, 1.0, 1e-8, 1e-10, [(0.5, 0.5)]),                                                                         # max_iterations = 0
    (2, 1.0, 1e-8, 1e-10, [(0.5, 0.5), (0.25, 0.25), "it", "it"]),                                               # counted past max + 1: "too many iterations"
]

DOMAIN_FILES = {
    "repo_domain.dat": None,            # modular-pnp_b200/host/domain.dat (the VALUES of benchmarks/pnp_exact_solution/domain.dat)
    "repo_domain_stokes.dat": None,     # modular-pnp_b200/host/domain_stokes.dat
    "comments.dat": "% a comment line\n[ section header ] with words\n| another style\nlength_x = 3.5 % trailing words are skipped\n"
                    "grid_x = 7\n%grid_y = 99\n  grid_z    =    12   \nmesh_output = ./out/m.pvd\n",
    "unknown_key.dat": "length_x = 2.0\nbogus_key = 5\ngrid_y = 4\nlength_q 7\ngrid_z = 6\n",
    "malformed_stops.dat": "length_x = 2.5\ngrid_x 9\ngrid_y = 33\nlength_y = 7.0\n",          # no '=' after grid_x: reading stops, the rest keeps the defaults
    "bad_number.dat": "grid_y = 5\ngrid_x = abc\ngrid_z = 8\n",
    "bad_real.dat": "length_y = 4.0\nlength_x = x1\nlength_z = 8\n",
    "int_from_real.dat": "grid_x = 12.7\ngrid_y = 6\n",                                         # %d reads 12, the rest of the line is skipped
    "grid_too_small.dat": "grid_x = 1\ngrid_y = 10\ngrid_z = 10\n",                             # range check: ERROR_INPUT_PAR
    "negative_length.dat": "length_z = -1.0\n",
    "negative_time.dat": "grid_time = -1\n",
    "strings.dat": "mesh_file = ./meshes/cube.xml\nsubdomain_file = sub.xml\nsurface_file = surf.xml\nmesh_output = none\nref_length = 1.0e-9\nlength_time = 2.5\ngrid_time = 3\n",
    "empty.dat": "",
    "last_line_no_newline.dat": "grid_x = 21\ngrid_y = 22\ngrid_z = 23",
    "value_missing_at_eof.dat": "grid_x = 21\ngrid_y =",
    "key_repeated.dat": "grid_x = 4\ngrid_x = 5\nlength_x = 1.5\nlength_x = 2.5\n",
}
DOMAIN_SPECIAL = ["<null>", "<missing file>"]

# (d[5], g[4], mesh_file)
BUILD_CASES = [
    ([1e-9, 2.0, 0.4, 0.4, 0.0], [50, 10, 10, 0], "none"),
    ([1.0, 10.0, 1.0, 1.0, 0.0], [50, 5, 5, 0], "none"),
    ([1.0, float("nan"), 1.0, 1.0, 0.0], [4, 4, 4, 0], "none"),
    ([1.0, 1.0, 1.0, float("nan"), 0.0], [4, 4, 4, 0], "none"),
    ([1.0, 1.0, 1.0, 1.0, 0.0], [4, 4, 4, 0], "mesh.xml"),
    ([1.0, 3.0, 2.0, 1.0, 0.0], [1, 2, 3, 0], "none"),          # domain_build itself does not look at the grid
]


# --- driver -----------------------------------------------------------------------------------------------------------
class _Capture:
    """what the C side prints on fd 1 during the block"""

    def __enter__(self):
        sys.stdout.flush()
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        self.saved = os.dup(1)
        os.dup2(self.tmp.fileno(), 1)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 1); os.close(self.saved)
        self.tmp.seek(0); self.text = self.tmp.read().decode(); self.tmp.close()


def _bind(lib, pre):
    f = lambda n: getattr(lib, pre + n)     # noqa: E731
    f("newton_new").restype = C.c_void_p
    f("newton_new").argtypes = [C.c_ulong, C.c_double, C.c_double, C.c_double]
    for n in ("newton_delete", "newton_update_iteration", "newton_print_status"):
        f(n).argtypes = [C.c_void_p]; f(n).restype = None
    f("newton_update_residuals").argtypes = [C.c_void_p, C.c_double, C.c_double]; f("newton_update_residuals").restype = None
    for n in ("newton_update_max_residual", "newton_update_rel_residual"):
        f(n).argtypes = [C.c_void_p, C.c_double]; f(n).restype = None
    for n in ("newton_needs_to_iterate", "newton_converged"):
        f(n).argtypes = [C.c_void_p]; f(n).restype = C.c_int
    f("newton_state").argtypes = [C.c_void_p, C.POINTER(C.c_double)]; f("newton_state").restype = None
    f("domain_param_input").argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p]
    f("domain_param_input").restype = C.c_int
    f("domain_build").argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.POINTER(C.c_double)]
    f("domain_build").restype = C.c_int
    return f


def _num(x):
    return None if x != x else x           # NaN -> null in the JSON


def run_cases(lib, pre, workdir):
    f = _bind(lib, pre)
    out = {"newton": [], "domain": {}, "build": []}
    for (mx, r0, rt, mt, seq) in NEWTON_CASES:
        h = f("newton_new")(mx, r0, rt, mt)
        st = (C.c_double * 3)()
        trace = []

        def snap():
            f("newton_state")(h, st)
            trace.append([f("newton_needs_to_iterate")(h), f("newton_converged")(h), int(st[0]), _num(st[1]), _num(st[2])])
        snap()
        k = 0
        while f("newton_needs_to_iterate")(h) and k < len(seq) and seq[k] != "it":
            f("newton_update_residuals")(h, seq[k][0], seq[k][1]); f("newton_update_iteration")(h); k += 1
            snap()
        for _ in [e for e in seq if e == "it"]:          # the counter alone, past the loop
            f("newton_update_iteration")(h)
            snap()
        with _Capture() as cap:
            f("newton_print_status")(h)
        # the two single-field updates as well (pnp_newton_solver.h:190 uses update_max_residual on its own)
        f("newton_update_max_residual")(h, 0.125); f("newton_update_rel_residual")(h, 0.25 * r0)
        snap()
        f("newton_delete")(h)
        out["newton"].append({"steps_taken": k, "trace": trace, "print_status": cap.text})
    names = list(DOMAIN_FILES) + DOMAIN_SPECIAL
    for name in names:
        if name == "<null>":
            path = None
        elif name == "<missing file>":
            path = os.path.join(workdir, "does_not_exist.dat").encode()
        elif DOMAIN_FILES[name] is None:
            path = os.path.join(HOST, name[len("repo_"):]).encode()
        else:
            p = os.path.join(workdir, name)
            with open(p, "w") as fh:
                fh.write(DOMAIN_FILES[name])
            path = p.encode()
        d = (C.c_double * 5)(); g = (C.c_int * 4)(); s = C.create_string_buffer(512)
        with _Capture() as cap:
            status = f("domain_param_input")(path, d, g, s)
            sys.stdout.flush()
        strs = [s.raw[128 * q:128 * (q + 1)].split(b"\0")[0].decode() for q in range(4)]
        text = cap.text.replace(workdir, "<tmp>")
        out["domain"][name] = {"status": status, "d": [_num(v) for v in d], "g": list(g), "strings": strs, "printed": text}
    for (d_, g_, mf) in BUILD_CASES:
        d = (C.c_double * 5)(*d_); g = (C.c_int * 4)(*g_); box = (C.c_double * 9)()
        with _Capture() as cap:
            built = f("domain_build")(d, g, mf.encode(), box)
        out["build"].append({"built": built, "box": [_num(v) for v in box], "printed": cap.text})
    return out


def _mirror_lib(tmp):
    so = os.path.join(tmp, "libhostmirror.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + HOST,
                           os.path.join(ROOT, "tests", "host_shims", "host_logic_shim.cpp"), "-o", so])
    return C.CDLL(so)


def _same(a, b, path=""):
    """exact comparison (these are decisions, integers, parsed literals and printed text), NaN == NaN"""
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], path + "/" + str(k))
    elif isinstance(a, list):
        assert isinstance(b, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, path + "/" + str(i))
    else:
        assert a == b, "%s: %r != %r" % (path, a, b)


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


def test_host_mirrors_equal_the_reference_golden(golden, tmp_path):
    got = json.loads(json.dumps(run_cases(_mirror_lib(str(tmp_path)), "mir_", str(tmp_path))))
    # the one deliberate difference: the reference prints its "Could not open file" line and then leaves through
    # fasp_chkerr; the mirror prints the same line and returns the same status
    _same(got, golden)


def test_golden_is_what_the_compiled_reference_says(golden, tmp_path):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libpnpref_host.so not built (needs /root/reference at build time)")
    ref = json.loads(json.dumps(run_cases(C.CDLL(REF_LIB), "ref_", str(tmp_path))))
    _same(ref, golden)


def test_cases_cover_the_decisions(golden):
    """the golden itself exercises every branch the mirrors restate"""
    fin = [c["trace"][-2] for c in golden["newton"]]
    assert any(t[1] == 1 for t in fin) and any(t[1] == 0 for t in fin)                     # converged and not converged
    assert any("too many iterations" in c["print_status"] for c in golden["newton"])
    assert any("max residual too large" in c["print_status"] for c in golden["newton"])
    st = {k: v["status"] for k, v in golden["domain"].items()}
    assert st["repo_domain.dat"] == 0 and st["<null>"] == 0 and st["<missing file>"] == -10
    assert st["grid_too_small.dat"] < 0 and st["negative_length.dat"] < 0 and st["malformed_stops.dat"] == 0
    assert golden["domain"]["repo_domain.dat"]["g"][:3] == [50, 10, 10] and golden["domain"]["repo_domain.dat"]["d"][1:4] == [2.0, 0.4, 0.4]
    assert golden["domain"]["malformed_stops.dat"]["g"][1] == 10                            # grid_y after the malformed entry: default
    assert [b["built"] for b in golden["build"]] == [1, 1, 0, 0, 0, 1]
    assert golden["build"][0]["box"] == [-1.0, -0.2, -0.2, 1.0, 0.2, 0.2, 50.0, 10.0, 10.0]


# --- the manufactured problem's data and the Dirichlet set (inputs of every parity test and of the bench) ----------------
def _ref_lib():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libpnpref_host.so not built (needs /root/reference at build time)")
    L = C.CDLL(REF_LIB)
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    i4 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    u1 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    L.ref_exact_problem_point.argtypes = [C.c_double, f8]
    L.ref_linear_function.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, f8, f8, C.c_int, f8, f8]
    L.ref_dirichlet_inside.argtypes = [C.c_int, i4, f8, f8, C.c_double, C.c_int, f8, u1, u1]
    return L


def _ref_problem(L, xyz):
    """the reference's own functions (benchmarks/pnp_exact_solution/main.cpp:24-78, :293-302 with src/dirichlet.cpp's
    Linear_Function) evaluated at the vertices: what main.cpp interpolates into its coefficient Functions"""
    X = xyz.reshape(-1, 3)
    nv = len(X)
    out = np.zeros((nv, 16))
    for v in range(nv):
        L.ref_exact_problem_point(float(X[v, 0]), out[v])
    lo, hi = np.zeros(16), np.zeros(16)
    L.ref_exact_problem_point(-1.0, lo); L.ref_exact_problem_point(+1.0, hi)
    xmin, xmax = float(X[:, 0].min()), float(X[:, 0].max())
    uu0 = np.zeros((nv, 3))
    for k in range(3):           # PDE::set_DirichletBC: one scalar Linear_Function per component, coordinate 0
        col = np.zeros(nv)
        L.ref_linear_function(0, xmin, xmax, 1, lo[k:k + 1].copy(), hi[k:k + 1].copy(), nv, np.ascontiguousarray(X.ravel()), col)
        uu0[:, k] = col
    return dict(exact=out[:, 0:3].ravel(), diff=out[:, 9:12].ravel(), fixed=out[:, 12].copy(), reac=out[:, 13:16].ravel(),
                uu0=uu0.ravel(), eps=np.ones(nv), val=np.tile([0.0, 1.0, -1.0], nv))


@pytest.mark.parametrize("grid,lengths", [((50, 10, 10), (2.0, 0.4, 0.4)), ((8, 8, 8), (2.0, 2.0, 2.0)), ((7, 3, 5), (2.0, 1.0, 1.6))])
def test_problem_data_equal_the_references_own_functions(oracle, grid, lengths):
    """The coefficient, source, exact-solution and initial-guess arrays fed to the oracle, to the device (tests) and to
    the bench equal what the reference's own functions give at the same vertices: bit for bit in the oracle's statement
    (oracle/assemble.cpp, orc_exact_problem), to one unit in the last place in the numpy statements
    (modular-pnp_b200/problems.py: bench, multi-GPU fields)."""
    from importlib import import_module
    from pnpb200_loader import load_package
    load_package()
    problems = import_module("modular_pnp_b200.problems")
    L = _ref_lib()
    P = oracle.Problem(*grid, *lengths)
    ref = _ref_problem(L, P.xyz)
    mine = dict(exact=P.exact, diff=P.diff, fixed=P.fixed, reac=P.reac, uu0=P.uu0, eps=P.eps, val=P.val)
    py = problems.exact_solution_problem(*grid, *lengths)
    pf = problems.exact_solution_fields(P.xyz[0::3])
    def one_ulp(a, b):
        return np.all(np.abs(a - b) <= np.spacing(np.abs(b)))
    for k in ref:
        assert np.array_equal(mine[k], ref[k]), "oracle: " + k          # same libm, same expression order: bit-equal
        # numpy's vectorised exp differs from libm's in the last bit at a few points (measured: `reac` at one of the 51
        # x-planes of the 50 x 10 x 10 box, 6e-17 relative; everything else bit-equal): bar = one unit in the last place
        assert one_ulp(py[k], ref[k]), "problems.exact_solution_problem: " + k
        if k != "uu0" or lengths[0] == 2.0:      # exact_solution_fields interpolates over the benchmark's x-range [-1, 1]
            assert one_ulp(pf[k], ref[k]), "problems.exact_solution_fields: " + k
    for k in ("exact", "diff", "fixed", "uu0", "eps", "val"):
        assert np.array_equal(py[k], ref[k]), k


def test_dirichlet_vertex_set_equals_the_references_subdomain(oracle):
    """src/pde.cpp:433-461 + src/dirichlet.cpp:24-37: Dirichlet_Subdomain({0}, mesh_min, mesh_max, 0.1 * rmin).inside on the
    boundary vertices = the vertex set the oracle (and through it every test) constrains."""
    L = _ref_lib()
    for grid, lengths in [((10, 4, 4), (2.0, 0.4, 0.4)), ((5, 6, 3), (2.0, 2.0, 1.0))]:
        P = oracle.Problem(*grid, *lengths)
        X = P.xyz.reshape(-1, 3)
        mn, mx = X.min(axis=0), X.max(axis=0)
        on_b = (np.isclose(X, mn) | np.isclose(X, mx)).any(axis=1).astype(np.uint8)
        h = np.array(lengths) / np.array(grid)
        # inradius of a Kuhn tetrahedron of an hx x hy x hz brick: 3 V / (sum of the face areas); V = hx hy hz / 6
        rmin = min(_kuhn_inradii(h))
        inside = np.zeros(P.nv, dtype=np.uint8)
        L.ref_dirichlet_inside(1, np.array([0], dtype=np.int32), np.ascontiguousarray(mn), np.ascontiguousarray(mx), 0.1 * rmin,
                               P.nv, np.ascontiguousarray(X.ravel()), on_b, inside)
        assert np.array_equal(inside, P.bcflag[0::3]) and 0 < inside.sum() < P.nv


def _kuhn_inradii(h):
    import itertools
    r = []
    for perm in itertools.permutations(range(3)):
        p = [np.zeros(3)]
        for d in perm:
            q = p[-1].copy(); q[d] += h[d]; p.append(q)
        p = np.array(p)
        vol = abs(np.linalg.det(p[1:] - p[0])) / 6.0
        area = 0.0
        for f in itertools.combinations(range(4), 3):
            a, b, c = p[list(f)]
            area += 0.5 * np.linalg.norm(np.cross(b - a, c - a))
        r.append(3.0 * vol / area)
    return r


REF_DIODE_LIB = os.path.join(ROOT, "oracle", "_ref", "libpnpref_diodephys.so")
GOLDEN_DIODE = os.path.join(ROOT, "tests", "golden", "diode_problem_points.json")
DIODE_X = [-1.0 + 2.0 * i / 40 for i in range(41)] + [-0.05, 0.05, -0.0499999, 0.0499999, -1e-12, 1e-12, 0.0123, -0.0377, 0.051, -0.051]
DIODE_VOLTS = [0.0, -0.3, 0.1, 0.5]


def _diode_ref_table(L):
    f8 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.ref_diode_point.argtypes = [C.c_double, C.c_double, f8]
    tab = np.zeros((len(DIODE_VOLTS), len(DIODE_X), 20))
    for a, volt in enumerate(DIODE_VOLTS):
        for i, x in enumerate(DIODE_X):
            L.ref_diode_point(float(x), float(volt), tab[a, i])
    return tab


def _check_diode_fields(tab):
    from importlib import import_module
    from pnpb200_loader import load_package
    load_package()
    pr = import_module("modular_pnp_b200.problems")
    xs = np.array(DIODE_X)
    for a, volt in enumerate(DIODE_VOLTS):
        ref = tab[a]
        f = pr.diode_fields(xs, volt)
        left, right = pr.diode_contacts(volt)
        assert np.array_equal(f["diff"].reshape(-1, 3), ref[:, 0:3]) and np.array_equal(f["eps"], ref[:, 3])
        assert np.array_equal(f["fixed"], ref[:, 4]) and np.array_equal(f["uu0"].reshape(-1, 3), ref[:, 6:9])
        assert np.array_equal(left, ref[0, 9:12]) and np.array_equal(right, ref[0, 12:15])
        assert np.array_equal(f["val"].reshape(-1, 3), ref[:, 15:18]) and np.all(f["reac"] == 0.0) and np.all(ref[:, 19] == 0.0)
        assert pr.DIODE_POISSON_SCALE == ref[0, 5] and pr.DIODE_BETA == ref[0, 18]


def test_diode_problem_data_equal_the_references_diode_h():
    """benchmarks/pnp_diode/diode.h compiled where it lies (oracle/_ref/libpnpref_diodephys.so): diffusivities, relative
    permittivity, fixed charge (doping profile with its junction ramp), contacts, the Initial_Guess expression, the
    Poisson scale — at 51 abscissae (both sides of every branch point) and 4 bias voltages, BIT-EQUAL to
    problems.diode_fields / diode_contacts, the inputs of the diode tests."""
    if not os.path.exists(REF_DIODE_LIB):
        pytest.skip("oracle/_ref/libpnpref_diodephys.so not built (needs /root/reference at build time)")
    _check_diode_fields(_diode_ref_table(C.CDLL(REF_DIODE_LIB)))


def test_diode_problem_data_equal_the_committed_reference_values():
    with open(GOLDEN_DIODE) as fh:
        g = json.load(fh)
    assert [float.fromhex(h) for h in g["x"]] == [float(x) for x in DIODE_X] and g["volts"] == DIODE_VOLTS
    tab = np.array([[[float.fromhex(h) for h in row] for row in per_volt] for per_volt in g["values20"]])
    _check_diode_fields(tab)


GOLDEN_POINTS = os.path.join(ROOT, "tests", "golden", "exact_problem_points.json")
POINTS_X = [-1.0 + 2.0 * i / 50 for i in range(51)] + [-0.987654321, -0.3333333333333333, 1e-9, 0.123456789, 0.77, 0.9999]


def test_oracle_problem_data_equal_the_committed_reference_values(oracle):
    """the same comparison against COMMITTED outputs of the reference's functions (hex floats, exact), so that it also
    runs where oracle/_ref is absent: main.cpp:24-78 at 57 abscissae, and the scalar Linear_Function on [-1, 1]"""
    with open(GOLDEN_POINTS) as fh:
        g = json.load(fh)
    xs = np.array([float.fromhex(h) for h in g["x"]])
    ref = np.array([[float.fromhex(h) for h in row] for row in g["values16"]])
    uu0_ref = np.array([[float.fromhex(h) for h in row] for row in g["uu0"]])
    nv = len(xs)
    xyz = np.zeros((nv, 3)); xyz[:, 0] = xs
    eps = np.zeros(nv); fixed = np.zeros(nv); diff = np.zeros(3 * nv); val = np.zeros(3 * nv); reac = np.zeros(3 * nv)
    uu0 = np.zeros(3 * nv); exact = np.zeros(3 * nv)
    oracle.lib().orc_exact_problem(nv, np.ascontiguousarray(xyz.ravel()), eps, fixed, diff, val, reac, uu0, exact.ctypes.data)
    assert np.array_equal(exact.reshape(nv, 3), ref[:, 0:3]) and np.array_equal(diff.reshape(nv, 3), ref[:, 9:12])
    assert np.array_equal(fixed, ref[:, 12]) and np.array_equal(reac.reshape(nv, 3), ref[:, 13:16])
    assert np.array_equal(uu0.reshape(nv, 3), uu0_ref)
    assert np.all(eps == 1.0) and np.array_equal(val.reshape(nv, 3), np.tile([0.0, 1.0, -1.0], (nv, 1)))


if __name__ == "__main__":
    if "--write-golden" in sys.argv:
        with tempfile.TemporaryDirectory() as t:
            res = run_cases(C.CDLL(REF_LIB), "ref_", t)
        with open(GOLDEN, "w") as fh:
            json.dump(res, fh, indent=1)
        print("wrote", GOLDEN)
        L = _ref_lib()
        xyz = np.zeros((len(POINTS_X), 3)); xyz[:, 0] = POINTS_X
        r = _ref_problem(L, xyz.ravel())
        v16 = np.zeros((len(POINTS_X), 16))
        for i, x in enumerate(POINTS_X):
            L.ref_exact_problem_point(float(x), v16[i])
        with open(GOLDEN_POINTS, "w") as fh:
            json.dump({"source": "benchmarks/pnp_exact_solution/main.cpp:24-78 and src/dirichlet.cpp:73-90 compiled where they lie (oracle/_ref/libpnpref_host.so)",
                       "layout16": "exact_solution(3), exact_derivative(3), exact_second(3), diffusivities(3), fixed(1), reactions(3)",
                       "x": [float(x).hex() for x in POINTS_X], "values16": [[float(v).hex() for v in row] for row in v16],
                       "uu0": [[float(v).hex() for v in row] for row in r["uu0"].reshape(-1, 3)]}, fh, indent=0)
        print("wrote", GOLDEN_POINTS)
        tab = _diode_ref_table(C.CDLL(REF_DIODE_LIB))
        with open(GOLDEN_DIODE, "w") as fh:
            json.dump({"source": "benchmarks/pnp_diode/diode.h compiled where it lies (oracle/_ref/libpnpref_diodephys.so, oracle/ref_wrappers_diodephys.inc)",
                       "layout20": "diffusivities(3), relative_permittivity, fixed, 1/permittivity_factor, Initial_Guess(3), left_contact(3), right_contact(3), valencies(3), thermodynamic_beta, sum(reactions)",
                       "x": [float(x).hex() for x in DIODE_X], "volts": DIODE_VOLTS,
                       "values20": [[[float(v).hex() for v in row] for row in per_volt] for per_volt in tab]}, fh, indent=0)
        print("wrote", GOLDEN_DIODE)
