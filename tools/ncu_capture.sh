#!/bin/bash
# ncu --set full captures of the dominant kernels (run under gpurun, 1 GPU), after the same command ran clean.
# usage: tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
python tools/profile_forward.py 256 2 > $OUT/fwd_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/fwd_plain.log; exit 1; }
COMMON="--set full --clock-control none --import-source on --kernel-name-base demangled"
# fused separable, Cout = 512 on CTA pairs (cta_group::2), dil 1: 2nd launch of the 2nd forward (= layer 9)
timeout 400 ncu $COMMON -k 'regex:sep_pair2_kernel<\(int\)1' -s 6 -c 1 -o $OUT/${TAG}_sep512 -f python tools/profile_forward.py 256 2 > $OUT/ncu_sep512.log 2>&1
# fused separable, first layer (32 -> 64 at 184x184, HBM-bound)
timeout 400 ncu $COMMON -k 'regex:sep_conv_kernel<\(int\)64' -s 1 -c 1 -o $OUT/${TAG}_sep32 -f python tools/profile_forward.py 256 2 > $OUT/ncu_sep32.log 2>&1
# dense 3x3 128->128 from halo patches (CTA-pair kernel), dilation 1 and 2: launches of the 2nd forward
timeout 400 ncu $COMMON -k 'regex:conv3x3_halo_kernel<\(int\)1' -s 9 -c 1 -o $OUT/${TAG}_conv3x3_d1 -f python tools/profile_forward.py 256 2 > $OUT/ncu_c3d1.log 2>&1
timeout 400 ncu $COMMON -k 'regex:conv3x3_halo_kernel<\(int\)2' -s 5 -c 1 -o $OUT/${TAG}_conv3x3_d2 -f python tools/profile_forward.py 256 2 > $OUT/ncu_c3d2.log 2>&1
# fused head pair (initial stage, hidden 512)
timeout 400 ncu $COMMON -k 'regex:head_pair_kernel' -s 2 -c 1 -o $OUT/${TAG}_heads -f python tools/profile_forward.py 256 2 > $OUT/ncu_heads.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
# decode chain on the forward's own maps (batch 256): 2nd decode pass
timeout 400 ncu $COMMON -k 'regex:peaks_kernel' -s 1 -c 1 -o $OUT/${TAG}_peaks -f python tools/profile_decode.py 256 > $OUT/ncu_peaks.log 2>&1
timeout 400 ncu $COMMON -k 'regex:limbs_kernel' -s 1 -c 1 -o $OUT/${TAG}_limbs -f python tools/profile_decode.py 256 > $OUT/ncu_limbs.log 2>&1
timeout 400 ncu $COMMON -k 'regex:group_fast_kernel' -s 1 -c 1 -o $OUT/${TAG}_group -f python tools/profile_decode.py 256 > $OUT/ncu_group.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep
# the remaining kernel classes: first conv, stride-2 depthwise, 128-wide 1x1 GEMM (CTA pair)
timeout 400 ncu $COMMON -k 'regex:conv1_bf16_kernel' -s 1 -c 1 -o $OUT/${TAG}_conv1 -f python tools/profile_forward.py 256 2 > $OUT/ncu_conv1.log 2>&1
timeout 400 ncu $COMMON -k 'regex:dw_tma_kernel' -s 2 -c 1 -o $OUT/${TAG}_dw -f python tools/profile_forward.py 256 2 > $OUT/ncu_dw.log 2>&1
timeout 400 ncu $COMMON -k 'regex:gemm_conv2sm_kernel<\(bool\)0' -s 8 -c 1 -o $OUT/${TAG}_gemm1x1 -f python tools/profile_forward.py 256 2 > $OUT/ncu_gemm1x1.log 2>&1
ls -la $OUT/${TAG}_*.ncu-rep | wc -l
