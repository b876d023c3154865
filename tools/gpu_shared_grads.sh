python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
python -m pytest tests -m gpu -q -k "gradient or shared or intensity or one_launch or geom" 2>&1 | tail -4
for a in "--fit two-kernel" "--fit two-kernel --shared-grads"; do python bench.py --config c5 $a --steps 10 --warmup 3 --no-e2e --no-eager --no-cpu | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['mode'], round(d['value'],2), round(d['ms_per_step'],3), 'loss kernel', round(d['roofline']['kernel_ms'],3))"; done
