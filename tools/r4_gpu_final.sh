#!/bin/bash
# the GPU calls of the last session of round 2 (element types added late), in the order they were made; each line is one gpurun call
# whose output went to gpurun_out/<tag>*.txt (summaries copied to profiles/, see profiles/README.md)
T=${1:-r4}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_prism.py -q -m gpu 2>&1 | tail -5 > gpurun_out/${T}b_prism.txt                      # FV1 prisms vs the oracle
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}c_gputest.txt                                              # full suite
python bench.py > gpurun_out/${T}c_bench.json 2> gpurun_out/${T}c_bench.err                                                 # contract bench line (traffic from the re-stamped capture)
python -m pytest tests/test_gpu_parity_fvcr.py -q -m gpu -k "quad or hex or errors" 2>&1 | tail -3 > gpurun_out/${T}d_fvcrq.txt   # FVCR quad / hex vs the oracle
python -m pytest tests/test_gpu_cavity.py -q -m gpu -s -k "fvcr" 2>&1 | grep -E "FVCR quads|passed|failed" > gpurun_out/${T}e_fvcrq_cavity.txt
python -m pytest tests/test_gpu_cavity.py -q -m gpu -s -k "extruded and prism" 2>&1 | grep -E "^extruded|passed|failed" > gpurun_out/${T}f_prism_cavity.txt
python -m pytest tests/test_gpu_boundary.py -q -m gpu -k "prism" 2>&1 | tail -2 > gpurun_out/${T}g_prism_bnd.txt           # boundary discs on prisms
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}h_gputest.txt                                              # full suite, final build
python -m pytest tests/test_golden.py -q -m gpu -k "prism or quad_fvcr or hex_fvcr" 2>&1 | tail -2 > gpurun_out/${T}i_golden.txt
# next session: NSB_CONFIGS=6,7,8,9 python tools/config_bench.py   (timings of the late element types, not measured yet)
echo done
