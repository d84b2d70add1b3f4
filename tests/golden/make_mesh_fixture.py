"""Writes tests/golden/mshr_sphere_mesh.npz from the one unstructured DOLFIN mesh the reference ships as DATA:
benchmarks/pnp_pb/mesh1.xml.gz (an mshr tetrahedralisation: box with a spherical inclusion, 6 542 vertices, 33 373 cells;
DOLFIN XML, written by the reference's own tool chain). Used as a real-world input for the unstructured path (pattern,
oracle assembly, recursive-coordinate-bisection partitioner). Run in the authoring container only (needs /root/reference):
    python tests/golden/make_mesh_fixture.py
"""
import gzip
import os
import re

import numpy as np

SRC = "/root/reference/benchmarks/pnp_pb/mesh1.xml.gz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mshr_sphere_mesh.npz")

txt = gzip.open(SRC, "rt").read()
v = re.findall(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)" z="([^"]+)"', txt)
c = re.findall(r'<tetrahedron index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)" v3="(\d+)"', txt)
assert [int(r[0]) for r in v] == list(range(len(v))) and [int(r[0]) for r in c] == list(range(len(c)))
xyz = np.array([[float(a) for a in r[1:]] for r in v])
cells = np.array([[int(a) for a in r[1:]] for r in c], dtype=np.int32)
np.savez_compressed(OUT, xyz=xyz, cells=cells, source=np.array(SRC.replace("/root/reference/", "")))
print(OUT, xyz.shape, cells.shape, os.path.getsize(OUT))
