"""every upwind x stabilisation branch of the FV1 path at solution level: the Re = 100 cavity on quadrilaterals solved on the device
against the Ghia table of the reference's DrivenCavityLinesEval:  python tools/cavity_ghia_branches.py [cells ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples")); sys.path.insert(0, ROOT)
import warnings
warnings.filterwarnings("ignore")
import cavity
from plugin_navierstokes_b200 import tools

sizes = [int(a) for a in sys.argv[1:]] or [32]
for cells in sizes:
    for stab in ("fields", "flow"):
        for upw in ("no", "full", "skewed", "lps", "positive"):
            t0 = time.time()
            try:
                disc, coords, conn, u, hist = cavity.solve(2, cells, re=100.0, verbose=False, upwind=upw, stab=stab)
                r = tools.DrivenCavityLinesEval(u.cpu().numpy(), coords, conn, 100)["Ghia"]
                print("Re  100  %3d^2 quads  %-8s upwind + %-6s  %2d iterations (defect x %.1e)  Ghia  u(0.5, y): max %.4f avg %.4f | v(x, 0.5): max %.4f avg %.4f  [%.1f s]"
                      % (cells, upw, stab.upper(), len(hist) - 1, hist[-1] / hist[0], r["vertical"]["max_diff"], r["vertical"]["average_diff"],
                         r["horizontal"]["max_diff"], r["horizontal"]["average_diff"], time.time() - t0), flush=True)
                disc.close()
            except Exception as ex:
                print("Re  100  %3d^2 quads  %-8s upwind + %-6s  FAILED: %s" % (cells, upw, stab.upper(), str(ex)[:120]), flush=True)
