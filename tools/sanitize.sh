#!/bin/bash
# compute-sanitizer passes over the kernels that changed in round 2 (run under gpurun; slow: small cases only).
#   memcheck : out-of-bounds / misaligned accesses
#   racecheck: shared-memory hazards (the named-barrier statistics pass of the igemm epilogue, the Griffin-Lim frame rounds,
#              the distributed-shared-memory exchange of the GroupNorm backward)
set -x
SEL='test_conv_epilogue_statistics_feed_groupnorm and bf16 and (16x16 or 8x8 or 1d_ragged or up_16x16)'
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" 2>&1 | tail -6
compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "$SEL" 2>&1 | tail -8
compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "groupnorm_silu_backward and bf16" 2>&1 | tail -6
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "griffinlim_kernel_matches_oracle or groupnorm_silu_backward" 2>&1 | tail -6
# later in round 2: the multi-block attention kernel (P in tensor memory), the one-launch operand repack, the GroupNorm
# backward that parks dv in the dx buffer
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "(attention_matches_reference_formula and bf16) or training_helper_kernels or fused_dropout" 2>&1 | tail -6
compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention_matches_reference_formula and bf16 and (508 or 300 or 256)" 2>&1 | tail -6
# end of round 2: the option paths (FiLM, causal mask, pooled resamplers, Fourier-embedded conditioning, cond_sample), the
# two-CTA Griffin-Lim kernel and the rewritten small training kernels
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "(unet_forward_matches and (film or causal or pool or condembed) and bf16) or (signal_conditioned and bf16) or causal_attention or resamplers or griffinlim_other_frame or training_helper" 2>&1 | tail -6
