#!/bin/bash
# Refresh, in ONE gpurun call, what bench.py and DESIGN quote from profiles/: the launch lists of a denoiser call and of a
# training step, the DRAM capture of the igemm launches (summarised on the box, so that the bench line that follows quotes
# it), and the bench line itself.  Everything lands in gpurun_out/; copy the files into profiles/ afterwards.
set -x
O=gpurun_out
NV="--nvtx --nvtx-include denoiser_call/"
ncu $NV --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_unet.log 2>&1
ncu $NV -k regex:igemm_sm100 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_igemm_dram_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_dram.log 2>&1
python tools/summarize_dram.py $O/r2_igemm_dram_unet_b256.csv $O/igemm_dram_traffic.json $O/r2_ops_unet_b256.names.txt > /dev/null && cp $O/igemm_dram_traffic.json profiles/igemm_dram_traffic.json
TQ_TRAIN_GRAPH=0 ncu --nvtx --nvtx-include "train_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_b64.csv python tools/profile_train.py > $O/ncu_train.log 2>&1
python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err
tail -c 600 $O/bench_r2_final.json
