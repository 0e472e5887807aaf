set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for L in 32 16 8 4 2 1; do python bench.py --steps 5 --warmup 3 --no-cpu --lanes $L 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=4096 lanes',d['config']['lanes_per_env'],'value %.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],'ms',d['ms_per_step'])"; done
for L in 8 4 2 1; do python bench.py --steps 5 --warmup 3 --no-cpu --envs 65536 --lanes $L 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=65536 lanes',d['config']['lanes_per_env'],'value %.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],'ms',d['ms_per_step'])"; done
