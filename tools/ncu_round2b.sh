#!/bin/bash
# Second ncu pass of round 2 (after tools/ncu_families.sh): the kernels that changed in the second half of the round -- the templated
# stem kernels, the batched re-pack, the 128 -> 128 layer on its new route (tapped GEMM) next to the marching kernel it left, and the
# 32 -> 32 marching forward whose MMA-issue rate bounds the step.  Only compact metric blocks leave the box.
set -u
O=gpurun_out
mkdir -p $O
FULL="--set full"
LIGHT="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__inst_executed.sum.per_cycle_active,smsp__inst_executed.sum"
cap() { set_=$1; shift; name=$1; shift; regex=$1; shift; skip=$1; shift
  ncu $set_ --clock-control none --kernel-name-base demangled -k "regex:$regex" -s $skip -c 1 -o $O/tmp_$name -f "$@" > $O/ncu_$name.log 2>&1
  rc=$?
  ncu -i $O/tmp_$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py $name >> $O/r02b_ncu.txt
  rm -f $O/tmp_$name.ncu-rep $O/ncu_$name.log
  echo "$name rc=$rc"; }
: > $O/r02b_ncu.txt
python tools/one_step.py > $O/plain.log 2>&1 || echo "plain one_step failed"
python tools/ab_step.py --steps 2 > $O/plain.log 2>&1 || echo "plain ab_step failed"
cap "$LIGHT" stem_fwd          'stem_fwd_mma_kernel'   1 python tools/one_step.py
cap "$LIGHT" stem_wgrad        'stem_wgrad_mma_kernel' 1 python tools/one_step.py
cap "$LIGHT" pack_batched_bulk 'pack_batched_kernel'   7 python tools/ab_step.py --steps 2
cap "$FULL" march_fwd_32_32    'conv_march_kernel'  1 python tools/one_layer.py 32 32 128 1 fwd
cap "$FULL" tapped_128_128     'conv_tapped_gemm'   1 python tools/one_layer.py 128 128 32 1 fwd
REHR_MARCH_MAX_WIDE=100000 bash -c "$(declare -f cap); O=$O; cap '$FULL' march_fwd_128_128 'conv_march_kernel' 1 python tools/one_layer.py 128 128 32 1 fwd"
wc -l $O/r02b_ncu.txt
