#!/bin/bash
# One `ncu --set full` row per KERNEL CLASS of the training step, exported to CSV on the box -- the
# .ncu-rep files stay there (gpurun returns at most 64 MiB).  Usage (one GPU, ~8 min):
#     gpurun --timeout 900 -- 'bash tools/run_ncu_classes.sh r02'
# then here:  python tools/ncu_summary.py --csv gpurun_out/ncu_<tag>_*.csv <tag>
# Each class: skip the first S matching launches (warm-up steps), capture C.  The command under ncu is one eager step of
# the default workload (posenet_bs64) after 3 warm-up steps; its printed numbers are NOT bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
T0=$(date +%s)
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-modes"
run() {   # name regex skip count [extra bench args]
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $name "; then return; fi
  timeout 170 ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -f -o /tmp/ncu_$name \
      $CMD "$@" > gpurun_out/ncu_${TAG}_$name.log 2>&1
  local rc=$?
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_$name.csv 2>/dev/null
  echo "[t+$(( $(date +%s) - T0 ))s] $name rc=$rc rows=$(( $(wc -l < gpurun_out/ncu_${TAG}_$name.csv) - 2 ))"
}
# ---- conv engines, bf16 mode (what the bench line runs): 4th warm-up/timed step ----
# (ncu matches the regex against the function's base name: no return type, namespace or template arguments)
run conv1cta   '^k_tc_conv$'      60 6
run conv2cta   '^k_tc_conv2$'     90 6
run convhalo   '^k_tc_conv_halo$' 36 3
run wgrad1     '^k_tc_wgrad$'     30 3
run wgrad2     '^k_tc_wgrad2$'    80 6
# ---- element-wise / reduction / small kernels ----
run bnapply    'k_bn_apply'          100 3
run bnbwdapply 'k_bn_bwd_apply'      100 3
run bnfin      'k_bn_finalize_accum' 100 2
run bnbwdfin   'k_bn_bwd_finalize_accum' 100 2
run stempool   '^k_stem_pool$'       3 1
run stempoolb  'k_stem_pool_bwd'     3 1
run stems2d    'k_stem_s2d'          3 1
run pack       'k_pack_weights'      3 1
run unpack     'k_unpack_wgrads'     3 1
run loss       'k_loss'              3 1
run smallgemm  'k_small_gemm'        30 3
run gap        'k_gap'               6 2
run adam       'k_adam'              3 1
# ---- strict tensor-core mode: the same engines on split operand planes, plus its un-fused BN-backward reduction ----
run split_conv1cta '^k_tc_conv$'  60 4 --precision tc_split
run split_conv2cta '^k_tc_conv2$' 90 4 --precision tc_split
run split_wgrad2   '^k_tc_wgrad2$' 80 4 --precision tc_split
run split_sums     'k_channel_sums'  60 3 --precision tc_split
run split_bnapply  'k_bn_apply'      100 2 --precision tc_split
# source-level stall samples of two conv launches (tools/ncu_stalls.py)
ncu -i /tmp/ncu_conv2cta.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_conv2cta_source.csv 2>/dev/null
ncu -i /tmp/ncu_convhalo.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_convhalo_source.csv 2>/dev/null
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from geomapnet_b200 import build
open("gpurun_out/ncu_sources_digest.txt", "w").write(build._digest()[:16])
PY
ls -la gpurun_out/ncu_${TAG}_*.csv | awk '{print $5, $9}'; du -sh gpurun_out
echo "[t+$(( $(date +%s) - T0 ))s] done"
