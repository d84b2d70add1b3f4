"""-m "not gpu": the repo's parameter files (modular-pnp_b200/host/*.dat: rewritten files that carry the reference's
VALUES) against the reference's OWN files, through the product's parsers (fasp_param_input / fasp_param_init,
fasp_ns_param_input / fasp_ns_param_init, csrc/fasp_host.cpp):
  * every numeric field of the resulting itsolver / AMG / ILU (and _ns) parameter structs equals
    tests/golden/param_files.json, which holds the structs parsed from the reference's files
    (benchmarks/pnp_exact_solution/bsr.dat, benchmarks/pnp_stokes/{bcsr,bsr,ns}.dat, benchmarks/pnp_diode/bsr.dat,
    benchmarks/pnp_refine_exact/bsr.dat) — written by `python tests/test_param_files.py --write-golden`;
  * where /root/reference is present (the authoring container), the real files are parsed again and must still give
    the golden (so the product's parser is exercised on the reference's own formatting: comment banners, alignment)."""
import ctypes as C
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "param_files.json")
REF = "/root/reference/benchmarks"
HOST = os.path.join(ROOT, "modular-pnp_b200", "host")

# reference file -> the repo file that carries its values
PLAIN = {"pnp_exact_solution/bsr.dat": "bsr.dat", "pnp_stokes/bsr.dat": "bsr_stokes.dat", "pnp_stokes/bcsr.dat": "bcsr.dat",
         "pnp_refine_exact/bsr.dat": "bsr.dat", "pnp_diode/bsr.dat": "bsr_diode.dat"}
NS = {"pnp_stokes/ns.dat": "ns.dat"}


def _fields(st, prefix=""):
    out = {}
    for name, typ in st._fields_:
        v = getattr(st, name)
        if isinstance(v, C.Structure):
            out.update(_fields(v, prefix + name + "."))
        elif isinstance(v, (int, float)):
            out[prefix + name] = v
    return out


def _parse_plain(capi, path):
    L = capi.lib()
    inp, it, amg, ilu = capi.input_param(), capi.itsolver_param(), capi.AMG_param(), capi.ILU_param()
    L.fasp_param_input(path.encode(), C.byref(inp))
    L.fasp_param_init(C.byref(inp), C.byref(it), C.byref(amg), C.byref(ilu), None)
    return {"itsolver": _fields(it), "amg": _fields(amg), "ilu": _fields(ilu)}


def _parse_ns(capi, path):
    L = capi.lib()
    ns_in, it, amg, ilu, sch = capi.input_ns_param(), capi.itsolver_ns_param(), capi.AMG_ns_param(), capi.ILU_param(), capi.Schwarz_param()
    L.fasp_ns_param_input(path.encode(), C.byref(ns_in))
    L.fasp_ns_param_init(C.byref(ns_in), C.byref(it), C.byref(amg), C.byref(ilu), C.byref(sch))
    return {"itsolver_ns": _fields(it), "amg_ns": _fields(amg), "ilu": _fields(ilu)}


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


# what the path reads from these files: differences in fields the solvers never look at (output paths, print levels of
# the reference's own logging) are not parameter differences
IGNORED = ("print_level", "output_type")


def _same(a, b, what):
    for block in a:
        for k, v in a[block].items():
            if k.split(".")[-1] in IGNORED:
                continue
            assert b[block][k] == v, "%s: %s.%s = %r, reference file gives %r" % (what, block, k, b[block][k], v)


def test_repo_parameter_files_carry_the_reference_values(pkg, golden):
    for ref_name, mine in list(PLAIN.items()) + list(NS.items()):
        if mine is None:
            continue
        parse = _parse_ns if ref_name in NS else _parse_plain
        got = json.loads(json.dumps(parse(pkg.capi, os.path.join(HOST, mine))))
        _same(golden[ref_name], got, "host/%s vs benchmarks/%s" % (mine, ref_name))


def test_golden_is_what_the_reference_files_parse_to(pkg, golden):
    if not os.path.isdir(REF):
        pytest.skip("/root/reference absent (GPU box): the committed golden stands in")
    for ref_name in list(PLAIN) + list(NS):
        parse = _parse_ns if ref_name in NS else _parse_plain
        got = json.loads(json.dumps(parse(pkg.capi, os.path.join(REF, ref_name))))
        assert got == golden[ref_name], ref_name


def test_stokes_driver_default_differs_from_the_benchmarks_file_only_in_two_unread_keys(pkg, golden):
    """host/pnp_stokes reads host/bsr.dat for the PNP block unless told otherwise; the benchmark's own file
    (benchmarks/pnp_stokes/bsr.dat = host/bsr_stokes.dat) differs in precond_type (not read: the block solver always
    uses its AMG on the PNP block) and itsolver_restart 100 vs 30 (not reached: the basis holds <= 31 vectors and the
    PNP block converges in < 30 iterations) — and in nothing else."""
    a = json.loads(json.dumps(_parse_plain(pkg.capi, os.path.join(HOST, "bsr.dat"))))
    b = golden["pnp_stokes/bsr.dat"]
    diff = {(blk, k) for blk in b for k in b[blk] if a[blk][k] != b[blk][k] and k.split(".")[-1] not in IGNORED}
    assert diff == {("itsolver", "precond_type"), ("itsolver", "restart")}


def test_diode_file_selects_ilu(golden):
    """benchmarks/pnp_diode/bsr.dat:24,34-38: precond_type 4 (ILU), ILU_lfil 2 — what pnpb200_set_preconditioner(h, 1, 2)
    and the diode tests use"""
    g = golden["pnp_diode/bsr.dat"]
    assert g["itsolver"]["precond_type"] == 4 and g["ilu"]["ILU_lfil"] == 2 and g["itsolver"]["itsolver_type"] == 6


if __name__ == "__main__" and "--write-golden" in sys.argv:
    from pnpb200_loader import load_package
    capi = load_package().capi
    res = {}
    for ref_name in list(PLAIN) + list(NS):
        parse = _parse_ns if ref_name in NS else _parse_plain
        res[ref_name] = parse(capi, os.path.join(REF, ref_name))
    with open(GOLDEN, "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print("wrote", GOLDEN)
