# round-2 single-GPU evidence: tests, default bench (with config legs), reference arm, ncu captures + launch list
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c5_n1.json 2> gpurun_out/${TAG}_bench_c5_n1.err || tail -20 gpurun_out/${TAG}_bench_c5_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
bash tools/r02_ncu.sh $TAG > gpurun_out/${TAG}_ncu_driver.log 2>&1; tail -3 gpurun_out/${TAG}_ncu_driver.log
