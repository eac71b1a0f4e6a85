"""gather-only timing of a few hex configs: python tools/qb_gather.py N"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_bench import run
from plugin_navierstokes_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
modes = [("gather", capi.SCATTER_GATHER)]
for upwind, stab in (("full", "fields"), ("lps", "fields"), ("lps", "flow")):
    run("hex", n, upwind, stab, modes)
