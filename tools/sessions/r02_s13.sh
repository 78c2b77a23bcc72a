#!/bin/bash
# r02 session 13 (8 GPUs): world-2/4/8 sharding + peer-exchange tests, memcheck of the peer exchange, bench at N = 8 and N = 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
timeout 900 python -m pytest tests/test_sharding_gloo.py -q -m gpu > gpurun_out/r02_pytest_gpu_n8.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_n8.txt
timeout 600 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r02_sanitizer_memcheck_peer.%p.txt \
    python -m pytest tests/test_sharding_gloo.py -q -m gpu -k "peer_exchange and 2-10-1000003" > gpurun_out/r02_sanitizer_memcheck_peer_pytest.txt 2>&1; echo "memcheck peer rc=$?"; tail -2 gpurun_out/r02_sanitizer_memcheck_peer_pytest.txt
grep -h "ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck_peer.*.txt | sort | uniq -c
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['exchange']); print(json.dumps(d['strong'])[:1500]); print(d['e2e'])"
done
