#!/bin/bash
# round 2 final, part 4: the driver's own command lines (smoke, bench N=1 with its step counts, reference arm)
mkdir -p gpurun_out/final
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_driver_n1.json 2> gpurun_out/final/bench_driver_n1.err ) 2>&1 | grep real
python -c "import json;d=json.load(open('gpurun_out/final/bench_driver_n1.json'));print('N1',d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks'],d['roofline']['frac'],d['roofline']['share_of_step'],d['cpu_baseline']['value'],d['dropin_api']['value'])"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err ) 2>&1 | grep real
cut -c1-600 gpurun_out/final/bench_reference_arm.json
