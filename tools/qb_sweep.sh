#!/bin/bash
# usage: tools/qb_sweep.sh N -> rows-kernel variant sweep (hex N^3, LPS + FIELDS, gather)
N=${1:-128}
for cfg in "8 5" "8 6" "4 5" "4 6" "4 8"; do
  set -- $cfg
  echo "CH=$1 MINB=$2"
  NSB_ROWS_CH=$1 NSB_ROWS_MINB=$2 python -c "
import sys; sys.path.insert(0,'.')
from tools.quick_bench import run
from plugin_navierstokes_b200 import capi
run('hex', $N, 'lps', 'fields', [('gather', capi.SCATTER_GATHER)])
" 2>&1 | tail -1
done
