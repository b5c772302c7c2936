timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/kbench.py 256 10 4 2>&1 | grep -E "MAGPHASE|COMPLEX"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['features_on_device_value'])"
