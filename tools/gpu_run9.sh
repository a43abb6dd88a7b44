mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -20
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done | head
python - <<'PY'
import torch, os
for i in range(torch.cuda.device_count()):
    p = torch.cuda.get_device_properties(i)
    print(i, p.name, getattr(p, 'pci_bus_id', None), getattr(p, 'pci_device_id', None), getattr(p, 'pci_domain_id', None))
print(os.sched_getaffinity(0))
print(open('/proc/self/status').read().split('Mems_allowed_list')[1][:20])
PY
lscpu | grep -i -E "numa|socket|model name" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
tail -5 gpurun_out/r2_bench_2gpu.err
python tools/show_bench.py gpurun_out/r2_bench_2gpu.json | grep -E "^value|multi_gpu|e2e"

