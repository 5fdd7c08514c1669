#!/bin/bash
# One gpurun call: ncu on one launch of every kernel family of the hot list (DRAM bytes, tensor-pipe %, L2 ...).  Only the compact
# metric blocks leave the box (gpurun_out/r02_ncu_families.txt); the .ncu-rep files are deleted (64 MiB pull limit).
set -u
O=gpurun_out
mkdir -p $O
FULL="--set full"
LIGHT="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg"
cap() { set_=$1; shift; name=$1; shift; regex=$1; shift; skip=$1; shift
  ncu $set_ --clock-control none --kernel-name-base demangled -k "regex:$regex" -s $skip -c 1 -o $O/tmp_$name -f "$@" > $O/ncu_$name.log 2>&1
  rc=$?
  ncu -i $O/tmp_$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py $name >> $O/r02_ncu_families.txt
  rm -f $O/tmp_$name.ncu-rep $O/ncu_$name.log
  echo "$name rc=$rc"; }
: > $O/r02_ncu_families.txt
for spec in "64 32 128 1 fwd" "32 64 128 1 dgrad" "64 32 128 1 wgrad" "32 32 128 1 wgrad" "32 64 128 2 fwd" "256 256 16 1 fwd" "256 256 16 1 wgrad"; do
  python tools/one_layer.py $spec > $O/plain.log 2>&1 || echo "plain $spec failed"
done
cap "$FULL" march_fwd_64_32   'conv_march_kernel'  1 python tools/one_layer.py 64 32 128 1 fwd
cap "$FULL" march_dgrad_32_64 'conv_march_kernel'  1 python tools/one_layer.py 32 64 128 1 dgrad
cap "$FULL" wgrad_march_64_32 'wgrad_march_kernel' 1 python tools/one_layer.py 64 32 128 1 wgrad
cap "$FULL" wgrad_march_32_32 'wgrad_march_kernel' 1 python tools/one_layer.py 32 32 128 1 wgrad
cap "$FULL" tapped_s2_32_64   'conv_tapped_gemm'   1 python tools/one_layer.py 32 64 128 2 fwd
cap "$FULL" tapped_256_256    'conv_tapped_gemm'   1 python tools/one_layer.py 256 256 16 1 fwd
cap "$FULL" wgrad_256_256     'conv_wgrad_kernel'  1 python tools/one_layer.py 256 256 16 1 wgrad
python tools/one_step.py > $O/plain.log 2>&1 || echo "plain one_step failed"
cap "$LIGHT" in_apply_fwd      'in_apply_kernel<\(int\)0>'  22 python tools/one_step.py
cap "$LIGHT" in_apply_bwd      'in_apply_kernel<\(int\)1>'  22 python tools/one_step.py
cap "$LIGHT" in_reduce_bwd     'in_reduce_kernel<\(int\)1>' 22 python tools/one_step.py
cap "$LIGHT" stem_fwd          'stem_fwd_mma_kernel'   1 python tools/one_step.py
cap "$LIGHT" stem_wgrad        'stem_wgrad_mma_kernel' 1 python tools/one_step.py
cap "$LIGHT" pointwise_bwd     'pointwise_bwd_kernel'  1 python tools/one_step.py
python tools/misc_ops_once.py > $O/plain.log 2>&1 || echo "plain misc failed"
for k in sw_accumulate_kernel sw_finalize_kernel blur1d_kernel rot90_kernel mean_stack_kernel fba_combine_kernel upsample_d_kernel pointwise_fwd_kernel ndhwc_to_ncdhw_kernel; do
  cap "$LIGHT" $k "$k" 1 python tools/misc_ops_once.py
done
wc -l $O/r02_ncu_families.txt; du -sh $O
