mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pk_scatter -s 3 -c 1 -f -o gpurun_out/r2_scatter_packed python tools/ncu_scatter.py packed > gpurun_out/ncu_packed.log 2>&1
tail -1 gpurun_out/ncu_packed.log
