python -m pytest tests -m gpu -x -q > gpurun_out/s10_tests.log 2>&1; tail -3 gpurun_out/s10_tests.log
python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; tail -2 gpurun_out/s10_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/s10_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_share'], d['vxc_dft']['ms_per_build'], d['cpu_baseline']['value'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s10_launches.csv python bench.py --steps 1 --warmup 1 --profile-mode > /dev/null 2>&1
for k in k_tgemm_ws k_fold_reg k_offdiag_mma; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/s10_$k -f python bench.py --steps 1 --warmup 1 --profile-mode > /dev/null 2>&1
done
ncu --set full --clock-control none -k regex:k_gemm -s 8 -c 4 -o gpurun_out/s10_k_gemm -f python bench.py --steps 1 --warmup 1 --profile-mode > /dev/null 2>&1
ls -la gpurun_out/s10_*
