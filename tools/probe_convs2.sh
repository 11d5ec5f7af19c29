export TQ_IGEMM_PROF=1
for args in "256 32 32 128 128 3 0 0 0 128 2" "256 32 32 128 128 3 0 0 0 128 1" "256 16 16 256 256 3 0 0 0 256 2"; do
  for probe in 3 11 14 8; do
    echo "== probe=$probe args=$args"
    TQ_IGEMM_PROBE=$probe python tools/conv_bench.py one $args 2>&1 | grep -E "prof|TF/s" | tail -n 2
  done
done
