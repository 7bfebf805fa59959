#!/bin/bash
# round 2 final measurements, part 2: ncu launch list + full captures of the top kernels (profiles/capture.sh)
bash profiles/capture.sh > gpurun_out/capture.log 2>&1
tail -5 gpurun_out/capture.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
