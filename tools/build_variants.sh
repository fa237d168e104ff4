#!/bin/bash
# experiment helper: build extra launch-shape variants of the native CLOUDS kernel
# (images/APP_CLOUDS.native_w<W>c<C>.cubin = W warps per CTA, >= C CTAs per SM)
cd /root/repo/shaderbox_b200/csrc
for wc in "$@"; do
  w=${wc%%:*}; c=${wc##*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false --prec-div=true --prec-sqrt=true --ftz=false \
    -I../include -I../include/sbx -I../../include -I. -DAPP_CLOUDS=1 -DSBX_WARPS_PER_CTA=$w -DSBX_MIN_CTAS_PER_SM=$c \
    -DSBX_APP_HEADER='"native/app_clouds_native.h"' -Xptxas -v -cubin -o ../images/APP_CLOUDS.native_w${w}c${c}.cubin native/native_tu.cu 2>&1 | grep "Used\|spill" | head -3
done
