#!/bin/bash
# Round-2 profile capture, ONE GPU (run under gpurun; everything lands in gpurun_out/, summaries are made afterwards on
# the build box with tools/summarize_launches.py, tools/summarize_dram.py and tools/ncu_full_summary.py).
set -x
O=gpurun_out
NV="--nvtx --nvtx-include denoiser_call/"
# 1. launch list of one latent-UNet denoiser call (batch 256), decoder (micro-batch 64) and the Griffin-Lim launch
ncu $NV --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_unet.log 2>&1
ncu $NV --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_decoder_b64.csv python tools/profile_call.py 64 decoder > $O/ncu_dec.log 2>&1
# 2. DRAM traffic of every igemm launch of that call (roofline.traffic of bench.py)
ncu $NV -k regex:igemm_sm100 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_igemm_dram_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_dram.log 2>&1
# 3. full sets: the fp64 Griffin-Lim launch, the GroupNorm launches of a call, two igemm launches (N = 256 and N = 128 tiles)
ncu $NV --set full --clock-control none --import-source on -k regex:griffinlim_fused -c 1 -o $O/r2_gl_fp64 -f python tools/profile_call.py 256 gl > $O/ncu_gl.log 2>&1
# (the reports of many launches are tens of MB: only their raw-page CSV travels back -- gpurun_out/ is capped at 64 MiB)
ncu $NV --set full --clock-control none -k regex:gn_apply -c 20 -o /tmp/r2_gn_apply -f python tools/profile_call.py 256 unet > $O/ncu_gn.log 2>&1
ncu -i /tmp/r2_gn_apply.ncu-rep --page raw --csv > $O/r2_gn_apply.raw.csv
ncu $NV --set full --clock-control none -k regex:igemm_sm100 -c 14 -o /tmp/r2_igemm -f python tools/profile_call.py 256 unet > $O/ncu_igemm.log 2>&1
ncu -i /tmp/r2_igemm.ncu-rep --page raw --csv > $O/r2_igemm.raw.csv
ncu -i $O/r2_gl_fp64.ncu-rep --page raw --csv > $O/r2_gl_fp64.raw.csv
ls -la $O | tail -20
