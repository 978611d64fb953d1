#!/bin/bash
# Round-2 evidence at HEAD, one gpurun call (one GPU).  Launch lists of one restore() step in both tensor-core precisions
# and ncu --set full captures of: the fused ResStack pair kernel (C = 64), the C = 128 ResStack convolutions (conv1 / conv2),
# the two-CTA pipeline for C = 128, a UNet 3x3 convolution at the top level (W = 127, C = 32), the GRU cluster kernel and the final-conv kernel.
# Numbers printed by anything under ncu are never bench values.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
for P in bf16 tf32; do
  VFX_PRECISION=$P $NCU --metrics gpu__time_duration.sum -s 0 -c 2000 --csv --log-file gpurun_out/r02_launches_$P.csv python tools/run_step.py 32 2 > gpurun_out/r02_launches_$P.log 2>&1
done
FULL="$NCU --set full --import-source on"
VFX_PRECISION=bf16 $FULL -k regex:resstack_pair -s 1 -c 2 -o gpurun_out/r02_pair_c64 -f python tools/bench_pair.py --dil 3,243 --iters 1 > gpurun_out/r02_ncu_pair.log 2>&1
VFX_PRECISION=bf16 $FULL -k regex:resstack_pair2 -s 1 -c 1 -o gpurun_out/r02_pair2_c128 -f python tools/bench_pair.py --impl 2 --C 128 --prec bf16 --dil 3 --iters 1 > gpurun_out/r02_ncu_pair2.log 2>&1
VFX_PRECISION=bf16 $FULL -k regex:conv_gemm_tc -s 2 -c 2 -o gpurun_out/r02_rs2 -f python tools/bench_conv.py --only 128 --B 32 --iters 1 --dil 3 > gpurun_out/r02_ncu_rs2.log 2>&1
VFX_PRECISION=bf16 $FULL -k regex:conv_gemm_tc -s 2 -c 2 -o gpurun_out/r02_unet_c32 -f python tools/bench_conv2d.py --C 32 --B 32 --iters 1 > gpurun_out/r02_ncu_unet.log 2>&1
VFX_PRECISION=bf16 $FULL -k regex:gru_cluster -s 0 -c 1 -o gpurun_out/r02_gru -f python tools/run_step.py 32 1 > gpurun_out/r02_ncu_gru.log 2>&1
VFX_PRECISION=bf16 $FULL -k regex:voc_post -s 0 -c 1 -o gpurun_out/r02_voc_post -f python tools/run_step.py 32 1 > gpurun_out/r02_ncu_post.log 2>&1
# tf32 C = 64 pair, fused on one SM with the residual stashed in TMEM (halo boxes d = 3, aligned boxes d = 243)
VFX_PRECISION=tf32 $FULL -k regex:resstack_pair3 -s 1 -c 2 -o gpurun_out/r02_pair3_tf32 -f python tools/bench_pair.py --prec tf32 --impl 3 --dil 3,243 --iters 1 > gpurun_out/r02_ncu_pair3.log 2>&1
# the two launches it replaced (VFX_FUSE_PAIR3=0): tf32 C = 64 pair in the encoded-stream form: conv1 (launch 1) and conv2 decode/encode (launch 3) of `--kind pairenc`
VFX_PRECISION=tf32 $FULL -k regex:conv_gemm_tc -s 1 -c 3 -o gpurun_out/r02_rs3_tf32 -f python tools/bench_conv.py --only 64 --B 32 --iters 1 --dil 3 --prec tf32 --kind pairenc > gpurun_out/r02_ncu_rs3_tf32.log 2>&1
ls -la gpurun_out/*.ncu-rep
