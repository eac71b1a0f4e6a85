"""lid-driven cavity solved on the device (examples/cavity.py) against the literature tables of the reference's
DrivenCavityLinesEval (incompressible/navier_stokes_tools.h:578-617):  python tools/cavity_ghia.py [Re] [cells ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples")); sys.path.insert(0, ROOT)
import warnings
warnings.filterwarnings("ignore")
import cavity
from plugin_navierstokes_b200 import tools

re = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
sizes = [int(a) for a in sys.argv[2:]] or [32, 64, 96]
elem = os.environ.get("NSB_CAVITY_ELEM", "quad")              # quad | tri (jittered by NSB_CAVITY_JITTER)
jit = float(os.environ.get("NSB_CAVITY_JITTER", "0"))
for cells in sizes:
    for upw in ("lps", "full"):
        t0 = time.time()
        if os.environ.get("NSB_CAVITY_DISC") == "fvcr":        # NavierStokesFVCR on triangles
            elem = "FVCR tri"
            disc, coords, conn, es, u, hist = cavity.solve_fvcr(cells, re=re, verbose=False, upwind=upw, jitter=jit)
            out = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, int(re), elem_sides=es)
        elif os.environ.get("NSB_CAVITY_DISC") in ("hex", "tet"):  # 3-D element types on the one-cell extrusion of the square
            elem = "extruded " + os.environ["NSB_CAVITY_DISC"]
            disc, coords, conn, u2d, hist = cavity.solve_extruded(os.environ["NSB_CAVITY_DISC"], cells, re=re, verbose=False, upwind=upw)
            out = tools.DrivenCavityLinesEval(u2d, coords, conn, int(re))
        else:
            disc, coords, conn, u, hist = cavity.solve(2, cells, re=re, verbose=False, upwind=upw, elem=elem, jitter=jit)
            out = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, int(re))
        for src, r in out.items():
            label = "%3d^2 %ss%s" % (cells, elem, " (jitter %.2f)" % jit if jit else "")
            print("Re %4d  %s  %-4s upwind%s  %2d iterations (defect x %.1e)  %-14s u(0.5, y): max %.4f avg %.4f | v(x, 0.5): max %.4f avg %.4f  [%.1f s]"
                  % (re, label, upw, "" if elem.startswith("FVCR") else " + FIELDS", len(hist) - 1, hist[-1] / hist[0], src, r["vertical"]["max_diff"], r["vertical"]["average_diff"],
                     r["horizontal"]["max_diff"], r["horizontal"]["average_diff"], time.time() - t0), flush=True)
        disc.close()
