set -x
python -m pytest tests/test_gpu_oneshot.py tests/test_gpu_dp2.py -m gpu -v --tb=short --timeout 600 -p no:cacheprovider > gpurun_out/r02_multi_gpu_tests.log 2>&1
tail -n 6 gpurun_out/r02_multi_gpu_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
done
for n in 1 2 4 8; do grep -o '"ms_per_step": [0-9.]*\|"e2e": {"value": [0-9.]*\|"value": [0-9.]*' gpurun_out/r02_scale_n$n.json | head -3; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 3 --no-e2e > gpurun_out/r02_cfg4_n8.json 2> gpurun_out/r02_cfg4_n8.err
grep -o '"ms_per_step": [0-9.]*\|"value": [0-9.]*' gpurun_out/r02_cfg4_n8.json | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/population_bench.py --candidates 256 --steps 50 --workers 4 > gpurun_out/r02_pop_n8.json 2> gpurun_out/r02_pop_n8.err
tail -n 1 gpurun_out/r02_pop_n8.json
