#!/bin/bash
# raw pinned-copy bandwidth on every GPU of the box at once (one process per GPU)
nproc; free -g | head -2 | tail -1; numactl -H 2>/dev/null | head -12; nvidia-smi topo -m 2>/dev/null | head -14
N=$(nvidia-smi -L | wc -l)
for i in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$i python tools/pcie_probe.py > /tmp/pcie_$i.log 2>&1 & done
wait
for i in $(seq 0 $((N-1))); do echo "gpu $i: $(cat /tmp/pcie_$i.log | tail -1)"; done
