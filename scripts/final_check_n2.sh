set -x
mkdir -p gpurun_out/r2f
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f/bench_ref_n2.json 2> gpurun_out/r2f/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2f/bench_n2.json 2> gpurun_out/r2f/bench_n2.err; echo "bench n2 rc=$?"
tail -c 3000 gpurun_out/r2f/bench_n2.json; tail -5 gpurun_out/r2f/bench_n2.err
