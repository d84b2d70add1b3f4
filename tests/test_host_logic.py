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
                           os.path.join(ROOT, "tests", "host_logic_shim.cpp"), "-o", so])
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


if __name__ == "__main__":
    if "--write-golden" in sys.argv:
        with tempfile.TemporaryDirectory() as t:
            res = run_cases(C.CDLL(REF_LIB), "ref_", t)
        with open(GOLDEN, "w") as fh:
            json.dump(res, fh, indent=1)
        print("wrote", GOLDEN)
