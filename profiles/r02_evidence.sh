#!/bin/bash
# one GPU call that gathers the round-2 evidence kept under profiles/: ncu --set full of the tensor-core kernel at the bench
# geometry, of the any-ratio kernel on config 5, of the biquad kernels; the launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python profiles/biquad_one.py > gpurun_out/r02_biquad_line.json 2> gpurun_out/r02_biquad.err
ncu --set full --clock-control none --import-source on -k regex:art_sinc_umma_kernel -s 6 -c 1 -o gpurun_out/r02_umma_final \
    python bench.py --steps 1 --warmup 1 --launches-per-step 4 --no-configs --no-cpu --no-e2e > gpurun_out/ncu_umma.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:art_sinc_generic -s 3 -c 1 -o gpurun_out/r02_generic_final \
    python profiles/asrc_one.py > gpurun_out/ncu_generic.log 2>&1
ncu --set full --clock-control none -k regex:bq_ -s 8 -c 8 -o gpurun_out/r02_biquad \
    python profiles/biquad_one.py > gpurun_out/ncu_biquad.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --launches-per-step 8 --no-configs --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python profiles/asrc_probe.py > gpurun_out/r02_asrc.jsonl 2>&1
tail -n 2 gpurun_out/r02_biquad_line.json; tail -n 4 gpurun_out/r02_asrc.jsonl
