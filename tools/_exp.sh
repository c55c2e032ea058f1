cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or bit_invariant" > gpurun_out/hs_test.log 2>&1; tail -4 gpurun_out/hs_test.log
for g in 5 6 7 8 10; do for b in 74 148; do echo "groups $g k3m $b"; SO100_GROUPS=$g SO100_K3M_BLOCKS=$b python tools/gpu_throughput.py 16384 100; done; done > gpurun_out/grp5.log 2>&1
grep -v "^groups" gpurun_out/grp5.log | awk '{print $3}' | paste - - | cat -n
python bench.py --no-extra > gpurun_out/bench11.log 2> gpurun_out/bench11.err; python -c "
import json; d=json.loads(open('gpurun_out/bench11.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
