set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; tail -2 gpurun_out/r02_gputest.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 600 gpurun_out/r02_bench.json
python profiles/microbench.py > gpurun_out/r02_microbench.txt 2>&1
python profiles/wave_probe.py > gpurun_out/r02_wave_probe.txt 2>&1
python profiles/config_probe.py > gpurun_out/r02_config_probe.txt 2>&1; cat gpurun_out/r02_config_probe.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-b3 > gpurun_out/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iif_conv_kernel -c 1 -f -o gpurun_out/r02_conv_wide python profiles/ncu_kernel_target.py conv 1184 1 > gpurun_out/r02_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:iif_product_kernel -c 1 -f -o gpurun_out/r02_prod_wide python profiles/ncu_kernel_target.py prod 592 1 > gpurun_out/r02_ncu_prod.log 2>&1
ls -la gpurun_out/*.ncu-rep
