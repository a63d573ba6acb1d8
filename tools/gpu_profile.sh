# Round-2 evidence run (B200, under gpurun): launch list of a bench run, `ncu --set full` captures of the dominant
# kernels condensed on the box (the reports themselves exceed what gpurun brings back), DRAM traffic for bench.py.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/p_bench_under_ncu.log 2>&1
prof() {  # name interp dtype fr
  R360_FRAMES=$4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 3 -c 1 -f -o /tmp/$1 \
      python tools/shape_sweep.py --interp $2 --dtype $3 --fr $4 --frames $5 --iters 1 > gpurun_out/p_ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep gpurun_out/$1.json > /dev/null 2>&1
  python tools/ncu_smem.py /tmp/$1.ncu-rep 12 > gpurun_out/$1.smem.txt 2>&1
  python tools/ncu_lines.py /tmp/$1.ncu-rep 30 > gpurun_out/$1.lines.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
prof r02_tiled_cubic_u8_b16 cubic u8 4 16
prof r02_tiled_linear_u8_b16 linear u8 4 16
prof r02_tiled_cubic_u16_b8 cubic u16 2 8
ls -la gpurun_out | head -30
