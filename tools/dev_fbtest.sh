for cfg in "96:384" "128:256" "160:192" "64:256" "96:256"; do
  b=${cfg%%:*}; t=${cfg##*:}
  MPC_FAST_BLOCKS=$b MPC_FAST_THREADS=$t python - <<PY
import os,sys
sys.path.insert(0,'.')
import torch
from tools.dev_sweep32 import run
# env must survive run()'s cleanup: re-inject
import tools.dev_sweep32 as d
for H in (50,17):
    d.run(H,4096,{"MPC_FAST_BLOCKS":"$b","MPC_FAST_THREADS":"$t"})
PY
done
