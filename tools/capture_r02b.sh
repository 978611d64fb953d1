#!/bin/bash
# Incremental round-2 evidence after conv_ts_tc.cu (tf32, C = 128, weights resident in TMEM) was added: the tf32 launch list of one
# restore() step again, and ncu --set full captures of the new kernel (conv1 form and conv2 encoded-stream form).  The other
# captures of tools/capture_r02.sh are unaffected (their kernels' sources did not change).
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
VFX_PRECISION=tf32 $NCU --metrics gpu__time_duration.sum -s 0 -c 2000 --csv --log-file gpurun_out/r02_launches_tf32.csv python tools/run_step.py 32 2 > gpurun_out/r02_launches_tf32.log 2>&1
VFX_PRECISION=tf32 $NCU --set full --import-source on -k regex:conv_ts -s 1 -c 2 -o gpurun_out/r02_ts_c128_tf32 -f python tools/bench_conv.py --only 128 --B 32 --iters 1 --dil 3 --prec tf32 --kind pairenc > gpurun_out/r02_ncu_ts.log 2>&1
ls -la gpurun_out/r02_ts_c128_tf32.ncu-rep
