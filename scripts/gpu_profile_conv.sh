#!/bin/bash
# ncu capture of the tcgen05 convolution kernel (tensor-pipe utilisation) inside the encoder forward
mkdir -p gpurun_out
timeout 300 python scripts/gpu_encoder_test.py 16 2048 > gpurun_out/encoder_test.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_fprop_tc_kernel -s 60 -c 20 -f -o gpurun_out/prof_conv_r01 \
    python scripts/gpu_encoder_test.py 16 2048 > gpurun_out/encoder_under_ncu.log 2>&1
