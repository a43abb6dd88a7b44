mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pk_scatter -s 1 -c 1 -f -o gpurun_out/r2_scatter_packed6 python tools/ncu_scatter.py packed6 > gpurun_out/ncu_packed6.log 2>&1
tail -2 gpurun_out/ncu_packed6.log
