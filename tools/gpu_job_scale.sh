#!/bin/bash
# multi-GPU run of the headline bench (one rank per GPU, torchrun): strong scaling (BASELINE config 3: global batch 32) at N = 8 and 2,
# weak scaling (32 prompts per GPU) at N = 8.  Use:  gpurun --gpus 8 -- bash tools/gpu_job_scale.sh
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
run() {  # N tag extra-args...
  local N=$1 tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-kernels --dropin 0 "$@" > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  echo "== $tag rc=$?"; tail -c 900 gpurun_out/scale_$tag.json; echo
  grep -E "NCCL INFO.*(nranks|NVLS|via P2P|Connected all)" gpurun_out/scale_$tag.err | head -4
}
run 8 strong_n8
run 8 weak_n8 --per-gpu-batch 32
run 2 strong_n2
