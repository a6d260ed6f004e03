run() { python bench.py --steps 20 --warmup 5 --no-extras "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f us  frac %.3f  %s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['kernel']))"; }
echo "rolled, unchained, pdl:"; run --no-chaining
echo "rolled, unchained, no pdl:"; MGB_PDL=0 run --no-chaining
echo "rolled, chained minb4:"; MGB_CHAIN_MINB=4 run
echo "rolled, chained minb6:"; MGB_CHAIN_MINB=6 run
for per in 224 112 128; do for st in 0 500 1000; do echo "per $per stagger $st pdl:"; MGB_STEP_PER=$per MGB_STAGGER_NS=$st run --no-chaining; done; done
echo "per 224 stagger 800 nopdl:"; MGB_PDL=0 MGB_STEP_PER=224 MGB_STAGGER_NS=800 run --no-chaining
echo "scalar:"; MGB_PACKED=0 run --no-chaining
