#!/usr/bin/env python
"""Generates tests/golden/fv1_fvcr_golden.npz (CASES) and tests/golden/elem_types_golden.npz (CASES_ELEM_TYPES: the element types
added at the end of round 2 -- FV1 prisms, FVCR quadrilaterals / hexahedra; a separate file so that the first one stays frozen).

The reference plugin cannot be compiled or run here (ugcore is absent), and it ships no golden
vectors, so these fixtures are FROZEN OUTPUTS OF THE CPU ORACLE (oracle/ns_oracle.c) on small seeded cases --
"parity unpinned": they pin the oracle against accidental drift and give the GPU tests a reference that does not
need the oracle at run time; they are not outputs of UG4.  Re-run only when the oracle is deliberately changed:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as ora          # noqa: E402
from plugin_navierstokes_b200 import meshgen   # noqa: E402

CASES = [
    # name, disc, elem, n, upwind, stab, extra flags
    ("cfg1_quad_full_fields_exact", "fv1", "quad", 5, "full", "fields", dict(exact_jac=1.0, kin_visc=1e-2)),
    ("cfg2_tri_lps_fields", "fv1", "tri", 5, "lps", "fields", dict(kin_visc=1e-3, peclet_blend=True)),
    ("cfg3_hex_lps_fields", "fv1", "hex", 3, "lps", "fields", dict(kin_visc=1e-2)),
    ("cfg5_hex_pos_flow_td", "fv1", "hex", 3, "positive", "flow", dict(kin_visc=1 / 1600, dt=1e-2, time_dependent=True)),
    ("tet_skewed_flow_pac", "fv1", "tet", 2, "skewed", "flow", dict(kin_visc=5e-3, pac=True, exact_jac=0.5)),
    ("cfg4_tet_fvcr_full", "fvcr", "tet", 2, "full", None, dict(kin_visc=1e-3, density=1.1)),
    ("tri_fvcr_lps_graddiv", "fvcr", "tri", 4, "lps", None, dict(kin_visc=1e-2, grad_div=0.2, laplace=True)),
]

CASES_ELEM_TYPES = [
    ("prism_lps_fields", "fv1", "prism", 3, "lps", "fields", dict(kin_visc=1e-2)),
    ("prism_pos_flow_td", "fv1", "prism", 2, "positive", "flow", dict(kin_visc=1 / 1600, dt=1e-2, time_dependent=True)),
    ("prism_full_flow_exact_peclet", "fv1", "prism", 2, "full", "flow", dict(kin_visc=5e-3, exact_jac=1.0, peclet_blend=True)),
    ("quad_fvcr_lps_graddiv", "fvcr", "quad", 4, "lps", None, dict(kin_visc=1e-2, grad_div=0.2, laplace=True)),
    ("hex_fvcr_full", "fvcr", "hex", 2, "full", None, dict(kin_visc=1e-3, density=1.1)),
]


def build(case):
    name, disc, elem, n, upwind, stab, flags = case
    coords, conn = meshgen.make_mesh(elem, n, jitter=0.2, seed=11)
    dim = coords.shape[1]
    E = ora.ELEM[elem]
    out = dict(coords=coords, conn=conn)
    if disc == "fv1":
        u = (meshgen.state_vortex3d if dim == 3 else meshgen.state_cavity2d)(coords, seed=12, noise=0.05)
        p = ora.make_params(elem=elem, upwind=upwind, stab=stab, **flags)
        rowptr, colind = ora.fv1_csr(E, conn, coords.shape[0])
        s0 = s1 = None
        if flags.get("time_dependent"):
            s0, s1 = u * 1.01 + 0.003, u * 0.97 - 0.002
            out.update(s0=s0, s1=s1)
        what = ora.JAC_A | ora.DEF_A
        vals, dfc = ora.assemble(p, conn, coords, u, rowptr, colind, what, sol0=s0, sol1=s1)
    else:
        es, n_side = meshgen.element_sides(elem, conn)
        rng = np.random.default_rng(13)
        u = rng.uniform(-1, 1, n_side * dim + conn.shape[0])
        p = ora.make_params(disc="fvcr", elem=elem, upwind=upwind, **flags)
        rowptr, colind = ora.fvcr_csr(E, es, n_side)
        vals, dfc = ora.assemble(p, conn, coords, u, rowptr, colind, ora.JAC_A | ora.DEF_A, elem_sides=es, n_side=n_side)
        out.update(elem_sides=es, n_side=np.int64(n_side))
    out.update(u=u, rowptr=rowptr, colind=colind, values=vals, defect=dfc)
    return {name + "/" + k: v for k, v in out.items()}


if __name__ == "__main__":
    # the first file is frozen since round 1: it is rewritten only on request (python make_golden.py --all)
    for fname, cases in (("fv1_fvcr_golden.npz", CASES if "--all" in sys.argv else []), ("elem_types_golden.npz", CASES_ELEM_TYPES)):
        if not cases:
            continue
        data = {}
        for c in cases:
            data.update(build(c))
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
        np.savez_compressed(path, **data)
        print("wrote", path, os.path.getsize(path), "bytes")
