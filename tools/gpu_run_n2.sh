mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/n2.err; echo "rc=$?" >> gpurun_out/n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n2.json 2>> gpurun_out/n2.err; echo "rc=$?" >> gpurun_out/n2.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/traffic_capture.py cfg2_seam_pole_views cfg4_u16_to_u16 cfg4_u16_to_f16 > gpurun_out/n2_traffic.log 2>&1
tail -n 4 gpurun_out/n2.err; cut -c1-600 gpurun_out/r02_bench_n2.json; cut -c1-300 gpurun_out/r02_bench_reference_n2.json; cut -c1-250 gpurun_out/n2_traffic.log
