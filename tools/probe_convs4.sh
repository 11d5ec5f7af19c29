# epilogue section profile (TQ_IGEMM_PROF=1) and clean timings (no profile) of the tile shapes that matter
for args in "256 32 32 128 128 3 1 1 1 128 2" "256 32 32 128 128 3 0 0 0 128 2" "256 16 16 256 256 3 1 1 1 256 2" "256 4 4 512 512 3 1 1 1 128 2"; do
  echo "== prof args=$args"
  TQ_IGEMM_PROF=1 python tools/conv_bench.py one $args 2>&1 | grep -E "prof|TF/s" | tail -n 3
done
for args in "256 32 32 128 128 3 1 1 1 128 2" "256 32 32 128 128 3 0 0 0 128 2" "256 32 32 128 128 3 1 1 1 128 1" "256 32 32 128 128 3 0 0 0 128 1" "256 32 32 384 128 3 0 0 1 128 2" "256 32 32 384 128 3 0 0 1 128 1" "256 16 16 256 256 3 1 1 1 256 2" "256 16 16 256 256 3 0 0 0 256 2" "256 16 16 256 256 3 0 0 0 256 1" "256 4 4 512 512 3 1 1 1 128 2" "256 4 4 512 512 3 1 1 1 256 2" "256 8 8 512 512 3 1 1 1 256 2"; do
  echo "== clean args=$args"
  python tools/conv_bench.py one $args 2>&1 | grep -E "TF/s" | tail -n 1
done
