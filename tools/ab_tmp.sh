run() { l=$1; shift; env "$@" timeout 100 python tools/ab_step.py --steps 40 --label "$l" 2>&1 | grep ms/step; }
run nodefer REHR_WGRAD_DEFER=0
run dfirst-32 A=1
run dfirst-64 REHR_DFIRST_MIN_VOXELS=524288
run dfirst-16 REHR_DFIRST_MIN_VOXELS=8192
run dfirst-0 REHR_DFIRST_MIN_VOXELS=0
run dfirst-32b A=1
