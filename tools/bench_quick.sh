#!/bin/bash
# short device-resident bench without roofline / cpu baseline; prints ms_per_step
python bench.py --steps 10 --warmup 3 --no-roofline --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step %.3f  graphs/s %.1f' % (d['ms_per_step'], d['value']))"
