#!/bin/bash
# ncu --set full captures of the three dominant forward kernels (run under gpurun, 1 GPU).
# usage: tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
python tools/profile_forward.py 256 2 > $OUT/fwd_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/fwd_plain.log; exit 1; }
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled"
# depthwise 512 channels (8th depthwise launch of the 2nd forward)
timeout 400 ncu $COMMON -k regex:dw_tma_kernel -s 21 -c 1 -o $OUT/${TAG}_dw512 -f python tools/profile_forward.py 256 2 > $OUT/ncu_dw.log 2>&1
# pointwise 512->512 (256-wide tiles: layers 5,6,7..11,20,22 -> 3rd launch of the 2nd forward is layer 7)
timeout 400 ncu $COMMON -k 'regex:gemm_conv_kernel<\(int\)256' -s 11 -c 1 -o $OUT/${TAG}_pw512 -f python tools/profile_forward.py 256 2 > $OUT/ncu_pw.log 2>&1
# dense 3x3 128->128 (im2col): first of the 2nd forward
timeout 400 ncu $COMMON -k 'regex:gemm_conv_kernel<\(int\)128, \(int\)128, \(bool\)1' -s 14 -c 1 -o $OUT/${TAG}_conv3x3 -f python tools/profile_forward.py 256 2 > $OUT/ncu_c3.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
