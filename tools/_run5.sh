timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_v15.json; python -c "import json; d=json.load(open('gpurun_out/bench_v15.json')); print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 120 --csv --log-file gpurun_out/launches_v15.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:irbpp -s 330 -c 2 -o gpurun_out/prof15 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | head
