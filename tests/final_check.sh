mkdir -p gpurun_out/final
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo smoke=$?
python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/final/pytest_gpu.log 2>&1; echo pytest=$?
tail -3 gpurun_out/final/pytest_gpu.log
python tests/run_configs.py c4 > gpurun_out/final/c4.jsonl 2> gpurun_out/final/c4.err; echo c4=$?
TC_VARIANTS="8,10" python tests/run_tc_gemm.py 2000 > gpurun_out/final/tc_gemm.log 2>&1
python bench.py --impl reference > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err; echo ref=$?
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo bench=$?
cat gpurun_out/final/c4.jsonl | cut -c1-400
cat gpurun_out/final/bench.json | cut -c1-300
tail -2 gpurun_out/final/smoke.log
