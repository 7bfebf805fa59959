#!/bin/bash
# strong scaling at N=8 on the final tree (global batch 32 -> 4 prompts per GPU)
mkdir -p gpurun_out/final
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/final/bench_n8.json 2> gpurun_out/final/bench_n8.err
python -c "import json;d=json.load(open('gpurun_out/final/bench_n8.json'));print('N8',d['value'],d['e2e']['value'],d['scaling'],d['config']['global_batch'],d['clocks'])"
grep -c "NCCL INFO" gpurun_out/final/bench_n8.err
