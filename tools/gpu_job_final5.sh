#!/bin/bash
# round 2 final, part 5 (2 GPUs): whole GPU suite once more after the last kernel changes, then the N=2 bench line
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final/bench_n2.json 2> gpurun_out/final/bench_n2.err
python -c "import json;d=json.load(open('gpurun_out/final/bench_n2.json'));print('N2',d['value'],d['e2e']['value'],d['scaling'],d['config']['global_batch'],d['clocks'])"
grep -c "NCCL INFO" gpurun_out/final/bench_n2.err
