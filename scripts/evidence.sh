# The commands behind profiles/r02_summary.md, run as ONE gpurun call (every ncu pass follows its plain run): bash scripts/evidence.sh
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_v37_bench_default.json 2> gpurun_out/r02_v37_bench_default.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/r02_v37_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v37_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hot_decode_kernel|sparse_decode_kernel|orbits_kernel|and_kernel|stream_kernel" -s 4 -c 4 -o gpurun_out/r02_v37_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity 0 > gpurun_out/ncu_full.log 2>&1
python bench.py --workload cfg1 --steps 5 --warmup 3 > gpurun_out/r02_v37_cfg1.json 2> gpurun_out/r02_v37_cfg1.err
python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r02_v37_cfg3.json 2> gpurun_out/r02_v37_cfg3.err
python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/r02_v37_cfg4_1gpu.json 2> gpurun_out/r02_v37_cfg4_1gpu.err
python scripts/batch_sweep.py > gpurun_out/r02_v37_batch_sweep.json 2> gpurun_out/r02_v37_batch_sweep.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_v37_reference_arm.json 2> gpurun_out/r02_v37_reference_arm.err
tail -c 300 gpurun_out/r02_v37_reference_arm.json; tail -3 gpurun_out/ncu_full.log
