#!/bin/bash
# Round 2, multi-GPU session (gpurun --gpus N): the multi-GPU parity tests at world N and the default bench line with
# its C5 / C3 blocks and the per-phase timings of the sharded assembly.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/r02_f_multi_gpu_tests_${N}gpu.txt 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r02_f_multi_gpu_tests_${N}gpu.txt | cut -c1-300 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f_bench_c2_${N}gpu.json 2> gpurun_out/r02_f_bench_c2_${N}gpu.err
tail -3 gpurun_out/r02_f_bench_c2_${N}gpu.err
python - <<PY
if [ "$1" == "ab" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-full-solve --no-extra --no-peer-memory > gpurun_out/r02_f_bench_c2_${N}gpu_nccl_sendrecv.json 2> gpurun_out/f_tmp.err
  python - <<PY2
import json
raw = open("gpurun_out/r02_f_bench_c2_${N}gpu_nccl_sendrecv.json").read()
d = json.loads([l for l in raw.splitlines() if l.startswith("{")][0])
print("c2 ncclSend/Recv", d["value"], d["phase_ms"], d.get("shard_assembly_ms"))
PY2
fi
import json
raw = open("gpurun_out/r02_f_bench_c2_${N}gpu.json").read()
d = json.loads([l for l in raw.splitlines() if l.startswith("{")][0])
print("c2", d["value"], d["phase_ms"], d.get("shard_assembly_ms"), "solve", d.get("solve_ms"), d.get("solve_iterations"), d.get("solved"))
print("c5", json.dumps(d.get("c5"))[:700])
print("c3", json.dumps(d.get("c3"))[:300])
PY
