#!/bin/bash
# One gpurun call (round 2, third pass): GPU test suite, the bench line, CUDA-event timings of the HBM-bound helper kernels
# (tools/hbm_kernels_bench.py) and ONE ncu run (light metric list, --clock-control none) over one launch sequence of the helper
# kernels.  Only compact text leaves the box.
set -u
O=gpurun_out
mkdir -p $O
timeout 240 python -m pytest tests -m gpu -x -q > $O/r02c_gputests.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/r02c_gputests.log
timeout 200 python bench.py > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench rc=$?"; head -c 400 $O/r02c_bench.json; echo
timeout 90 python tools/hbm_kernels_bench.py $O/r02c_hbm_kernels.json > $O/r02c_hbm_kernels.txt 2>&1; echo "hbm bench rc=$?"; cat $O/r02c_hbm_kernels.txt
LIGHT="--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 150 ncu $LIGHT --clock-control none --kernel-name-base demangled \
  -k "regex:sw_accumulate|sw_finalize|blur1d|mean_stack|fba_combine" -c 6 -o $O/tmp_r02c -f python tools/misc_ops_once.py > $O/ncu_r02c.log 2>&1
echo "ncu rc=$?"
ncu -i $O/tmp_r02c.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_extract.py helpers_r02c > $O/r02c_ncu_helpers.txt
rm -f $O/tmp_r02c.ncu-rep
wc -l $O/r02c_ncu_helpers.txt
