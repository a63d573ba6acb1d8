#!/bin/bash
# host topology of the box + the host-streamed leg of the config-3 sweep with and without CPU/NUMA binding
G=${1:-4}
mkdir -p gpurun_out
{ nproc; lscpu | grep -i -E "numa|socket|model name"; nvidia-smi topo -m; } > gpurun_out/numa_topo.txt 2>&1
for B in 0 1; do
  R360_NUMA_BIND=$B python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2954$B tools/config3_sweep.py --frames 96 --stream-frames 48 2>gpurun_out/numa_$B.err | tee -a gpurun_out/numa_probe.jsonl | cut -c1-600
done
head -40 gpurun_out/numa_topo.txt
