# Which resource bounds the igemm mainloop?  TQ_IGEMM_PROBE bit mask: 1 = no TMA operand loads, 2 = no MMAs,
# 4 = no epilogue work.  Results are garbage by construction; only the timing / cycle profile matters.
export TQ_IGEMM_PROF=1
for args in "256 32 32 128 128 3 0 0 0 128 2" "256 32 32 128 128 3 0 0 0 128 1" "256 32 32 128 128 3 0 0 0 64 2" "256 16 16 256 256 3 0 0 0 256 2" "256 16 16 256 256 3 0 0 0 256 1" "256 4 4 512 512 3 0 0 0 128 2"; do
  for probe in ${PROBES:-0 4 5 6 1 2 3}; do
    echo "== probe=$probe args=$args"
    TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "prof|TF/s" | tail -n 2
  done
done
