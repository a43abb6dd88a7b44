N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "rc=$?"
tail -3 gpurun_out/r02_bench_${N}gpu.err
python tools/show_bench.py gpurun_out/r02_bench_${N}gpu.json | grep -E "^value|multi_gpu|strong"
