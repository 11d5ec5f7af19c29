set -x
O=gpurun_out
NV="--nvtx --nvtx-include denoiser_call/"
ncu $NV --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_unet.log 2>&1
ncu $NV -k regex:igemm_sm100 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_igemm_dram_unet_b256.csv python tools/profile_call.py 256 unet > $O/ncu_dram.log 2>&1
TQ_TRAIN_GRAPH=0 ncu --nvtx --nvtx-include "train_step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_b64.csv python tools/profile_train.py > $O/ncu_train.log 2>&1
python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err
tail -c 600 $O/bench_r2_final.json
