set -x
mkdir -p gpurun_out/final
python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
for c in C1 C2 C4 C5; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/final/bench_$c.json 2> gpurun_out/final/bench_$c.err; done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 1 --warmup 3 --no-extra --no-e2e --no-cpu > gpurun_out/final/ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final/smoke.log
# full ncu sections for the hot kernels at a size where one replayed launch stays short (n = 20 000 x 2 Mb)
ncu --set full --clock-control none --import-source on -k regex:'k_pack4|k_sweep|k_block_n|k_block_d|k_cand_trim|k_slice' -c 12 -o gpurun_out/final/hot_20k python bench.py --n 20000 --steps 1 --warmup 0 --no-extra --no-e2e --no-cpu > gpurun_out/final/ncu_full.log 2>&1
