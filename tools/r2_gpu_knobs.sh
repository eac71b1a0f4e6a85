#!/bin/bash
# knob sweeps of the secondary configurations (config 1: general path, config 5: dense kernel)
for fm in 2 3 4; do for wpb in 1 2 3; do echo "config1 NSB_FLUX_MINB=$fm NSB_ROWS_WPB=$wpb"; NSB_CONFIGS=1 NSB_FLUX_MINB=$fm NSB_ROWS_WPB=$wpb python tools/config_bench.py 2>&1 | tail -1; done; done
for w in 1 2 4; do echo "config5 NSB_DENSE_WPB=$w"; NSB_CONFIGS=5 NSB_DENSE_WPB=$w python tools/config_bench.py 2>&1 | tail -1; done
