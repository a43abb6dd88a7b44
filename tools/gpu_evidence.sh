mkdir -p gpurun_out
P=gpurun_out/r02
timeout 1200 python bench.py --steps 10 --warmup 3 > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file ${P}_launches.csv python bench.py --steps 2 --warmup 3 --legs configs --extra-steps 3 --no-cpu-baseline --no-host-emit --no-pipelined --no-from-source > ${P}_launches_bench.json 2> ${P}_launches.err; echo "launches rc=$?"
python tools/ncu_step.py > ${P}_step_plain.log 2>&1; grep -E "PASS|LAUNCHES" ${P}_step_plain.log
SKIP=$(grep "PASS 1" ${P}_step_plain.log | awk '{print $3}')
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ -s ${SKIP} -c 80 -f -o ${P}_full python tools/ncu_step.py > ${P}_full.log 2>&1; echo "ncu full rc=$?"
tail -2 ${P}_full.log
python tools/show_bench.py ${P}_bench.json
ls -la gpurun_out/r02_*
