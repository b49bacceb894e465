set -x
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r2f/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r2f/bench_ref_n1.json 2> gpurun_out/r2f/bench_ref_n1.err; echo "ref rc=$?"
timeout 600 python bench.py --dtype bf16 --no-train --no-cpu-baseline > gpurun_out/r2f/bench_n1_bf16.json 2> gpurun_out/r2f/bench_n1_bf16.err; echo "bench bf16 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r2f/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python scripts/probe_train.py 512 16 bf16 prof cl > gpurun_out/r2f/train_prof.log 2>&1; echo "train prof rc=$?"
du -sh gpurun_out; ls -la gpurun_out/r2f
