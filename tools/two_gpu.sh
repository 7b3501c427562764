python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or c5" 2>&1 | tail -3 > gpurun_out/g2_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/g2_ref.json 2>> gpurun_out/g2_bench.err
cat gpurun_out/g2_pytest.log; cat gpurun_out/g2_bench.json; cat gpurun_out/g2_ref.json; tail -3 gpurun_out/g2_bench.err
