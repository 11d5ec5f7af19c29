# clean timings (no profile): full kernel vs MMA-only (probe 1: no TMA loads) vs TMA-only (probe 2: no MMAs)
for args in "256 32 32 128 128 3 0 0 0 128 2" "256 32 32 384 128 3 0 0 0 128 2" "256 16 16 256 256 3 0 0 0 256 2" "256 8 8 512 512 3 0 0 0 256 2" "256 32 32 128 8 3 0 0 0 64 2"; do
  for probe in 0 1 2 5 6; do
    echo "== probe=$probe args=$args"
    TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "TF/s" | tail -n 1
  done
done
