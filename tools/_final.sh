cd /root/repo
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_w.log 2>&1; tail -3 gpurun_out/pytest_w.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke7.log 2>&1; tail -2 gpurun_out/smoke7.log
python bench.py > gpurun_out/bench10.log 2> gpurun_out/bench10.err; tail -c 600 gpurun_out/bench10.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/benchref7.log 2>&1; tail -c 300 gpurun_out/benchref7.log
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
ncu --metrics $M --clock-control none -s 3100 -c 600 --csv --log-file gpurun_out/launches_r01h.csv python bench.py --steps 3 --warmup 5 --no-extra > gpurun_out/ncu_h.log 2>&1; tail -2 gpurun_out/ncu_h.log | cut -c1-300
SO100_GROUPS=1 SO100_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:phase_solve_heavy -s 400 -c 2 -o gpurun_out/r01k_medium python tools/gpu_throughput.py 16384 5 > gpurun_out/ncu_k.log 2>&1; tail -2 gpurun_out/ncu_k.log | cut -c1-300
