set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s4_bench.json')); print(d['value'], d['e2e'], d['roofline']['traced_gcups'], d['roofline']['recurrence_frac'])"
ncu --set full --clock-control none --import-source on -k regex:k_cluster_merge -s 1 -c 1 -o gpurun_out/kcluster python tools/run_tree.py 2000 400 0 > gpurun_out/kcluster.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_build_rows_t -s 1 -c 1 -o gpurun_out/krowst2 python tools/run_profile_batch.py 120 400 200 > gpurun_out/krowst2.log 2>&1
ls -la gpurun_out | tail -5
