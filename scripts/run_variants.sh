# per-step time of the quadrotor step kernel variants at the BASELINE shape (65 536 envs), graph-replayed; prints one line each
run() { python bench.py --steps 20 --warmup 5 --no-extras "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f us  frac %.3f  %s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['kernel']))"; }
echo "default (scalar, one CTA per SM, PDL):"; run
echo "scalar, PDL off:"; MGB_PDL=0 run
echo "scalar, 64-env CTAs:"; MGB_WIDE_KERNEL=0 run
echo "packed (MGB_PACKED=1):"; MGB_PACKED=1 run
