#!/bin/bash
# Which resource bounds the igemm kernel on a given shape?  One script for the probes of tools/conv_bench.py:
#
#   tools/probe_convs.sh [-p "0 1 2"] [-c] [-s "N H W cin cout k res emb stats bn cg" ...]
#
#   -p  TQ_IGEMM_PROBE bit masks to run (default "0"): 1 = no TMA operand loads, 2 = no MMAs, 4 = epilogue only hands the
#       accumulator back (results are garbage by construction; only the timing matters)
#   -c  clean timing: WITHOUT the in-kernel cycle profile (TQ_IGEMM_PROF=1 adds clock reads to every wait)
#   -s  a shape (repeatable); default = the shapes that matter: the 32x32 / 16x16 / 8x8 / 4x4 levels of the latent UNet, the
#       64- / 128- / 256-channel k = 5 levels of the 1D UNet at batch 64, the 128-wide 3x3 shapes of the pixel UNet / decoder
PROBES="0"
PROF=1
SHAPES=()
while getopts "p:cs:" o; do
  case $o in
    p) PROBES="$OPTARG" ;;
    c) PROF=0 ;;
    s) SHAPES+=("$OPTARG") ;;
    *) exit 2 ;;
  esac
done
if [ ${#SHAPES[@]} -eq 0 ]; then
  SHAPES=("256 32 32 128 128 3 1 1 1 0 0" "256 32 32 384 128 3 0 0 1 0 0" "256 16 16 256 256 3 1 1 1 0 0" "256 8 8 512 512 3 1 1 1 0 0"
          "256 4 4 512 512 3 1 1 1 0 0" "64 1 4064 64 64 5 1 0 1 0 0" "64 1 2032 128 128 5 1 0 1 0 0" "64 1 1016 256 256 5 1 0 1 0 0"
          "64 1 508 256 256 5 1 0 1 0 0" "16 128 128 128 128 3 1 0 1 0 0" "64 128 128 64 64 3 1 0 1 0 0")
fi
for args in "${SHAPES[@]}"; do
  for probe in $PROBES; do
    echo "== probe=$probe prof=$PROF args=$args"
    if [ "$PROF" = 1 ]; then
      TQ_IGEMM_PROF=1 TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "prof|TF/s" | tail -n 3
    else
      TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "TF/s" | tail -n 1
    fi
  done
done
