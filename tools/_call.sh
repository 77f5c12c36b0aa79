timeout 70 python bench.py --cpu-seconds 4 > gpurun_out/r1s_bench.json 2> gpurun_out/r1s_bench.err
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1s_launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 0.2 > /dev/null 2>&1
timeout 50 ncu --set full --clock-control none --import-source on -k regex:mol_rhs_tiled -s 3 -c 2 -o gpurun_out/r1s_tiled_full python bench.py --steps 3 --warmup 3 --cpu-seconds 0.2 > /dev/null 2>&1
timeout 40 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:mol_ -c 14 --csv --log-file gpurun_out/r1s_rk_launches.csv python tools/rk_bench.py 4096 tsit5 > /dev/null 2>&1
cat gpurun_out/r1s_bench.json
