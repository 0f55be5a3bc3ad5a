timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/bench_fkm.py 2>&1 | grep -o '"ms": [0-9.]*'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fkKernel -s 5 -c 1 -o gpurun_out/fkm_prof3 -f python tools/bench_fkm.py --steps 2 > gpurun_out/ncu_fkm3.log 2>&1
