# probes WITHOUT the cycle profile (clock reads perturb the loops): only the launch time matters
for args in "256 32 32 128 128 3 0 0 0 128 2" "256 32 32 128 128 3 0 0 0 128 1" "256 16 16 256 256 3 0 0 0 256 2" "256 32 32 128 128 3 0 0 0 64 1"; do
  for probe in 0 4 6 7 3; do
    echo "== probe=$probe args=$args"
    TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "TF/s" | tail -n 1
  done
done
