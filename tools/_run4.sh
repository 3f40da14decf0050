timeout 500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 150 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
IRBPP_LIB=$PWD/ir-bpp_b200/lib/libirbpp_fine.so timeout 100 python tools/_fine.py 2>&1 | tail -2
