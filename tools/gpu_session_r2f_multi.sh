#!/bin/bash
# Round 2, multi-GPU session (gpurun --gpus N): the multi-GPU parity tests at every world size the box allows and the
# default bench line with its C5 / C3 blocks and the per-phase timings of the sharded assembly; "ab": the same C2 step
# with the exchange through ncclSend / ncclRecv instead of peer memory.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_i_multi_gpu_tests_${N}gpu.txt 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_i_multi_gpu_tests_${N}gpu.txt | cut -c1-300 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_i_bench_c2_${N}gpu.json 2> gpurun_out/r02_i_bench_c2_${N}gpu.err
tail -3 gpurun_out/r02_i_bench_c2_${N}gpu.err | cut -c1-300
python tools/print_bench_line.py gpurun_out/r02_i_bench_c2_${N}gpu.json
if [ "$1" == "ab" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra --no-peer-memory > gpurun_out/r02_i_bench_c2_${N}gpu_nccl_sendrecv.json 2> gpurun_out/i_tmp.err
  python tools/print_bench_line.py gpurun_out/r02_i_bench_c2_${N}gpu_nccl_sendrecv.json
fi
