cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['stage_ms_per_step'])"
BENCH_DENSE=1 BENCH_SKIP_CPU=1 BENCH_SKIP_EXTRA=1 python bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dense', d['value'], d['ms_per_step'], d['config']['stage_ms_per_step'])"
