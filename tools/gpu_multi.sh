#!/bin/bash
# multi-GPU call: the C++ multi-rank driver tests + bench.py under torchrun on N GPUs ($1, default 2)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multirank.py tests/test_sharding.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/m${N}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/m${N}_pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?" >> gpurun_out/m${N}_bench.err
tail -3 gpurun_out/m${N}_pytest.log; tail -3 gpurun_out/m${N}_bench.err
