#!/bin/bash
# usage: tools/qb_env.sh N "ENV1=.. ENV2=.." ["..." ...] -> hex N^3 LPS+FIELDS gather timing under each environment
N=$1; shift
for envs in "$@"; do
  echo "[$envs]"
  env $envs python -c "
import sys; sys.path.insert(0,'.')
from tools.quick_bench import run
from plugin_navierstokes_b200 import capi
run('hex', $N, 'lps', 'fields', [('gather', capi.SCATTER_GATHER)])
" 2>&1 | tail -1
done
