timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 150 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_v14.json; cat gpurun_out/bench_v14.json | cut -c1-2600
timeout 150 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
timeout 200 python tools/e2e_probe.py 300 2>&1 | tail -10
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 120 --csv --log-file gpurun_out/launches_v14.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:irbpp -s 330 -c 2 -o gpurun_out/prof14 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/
