for S in 25 50 100 200 400; do
  python bench.py --no-cpu-baseline --solver-steps $S 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['solver_steps'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])"
done
