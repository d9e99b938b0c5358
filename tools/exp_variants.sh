for v in "" _mb3 _sk2 _sk2mb3; do
  export PGPU_LIB=$PWD/praline_b200/libpraline_b200$v.so
  echo "== variant '$v'"
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "paired_resident or packed_int16" 2>&1 | tail -2
  python tools/run_c3.py 3000 8 2>&1 | head -1 | cut -c1-220
  python tools/run_c3.py 1000 8 2>&1 | head -1 | cut -c1-220
done
