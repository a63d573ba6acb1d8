mkdir -p gpurun_out
tools/pipe_probe > gpurun_out/g_pipe_probe.jsonl 2>&1
export R360_MULTI_PCT=50
prof() {  # name library interp fr
  R360_LIBRARY=$2 R360_FRAMES=$4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:remap_tiled -s 3 -c 1 -f -o /tmp/$1 python tools/shape_sweep.py --interp $3 --fr $4 --iters 1 > gpurun_out/g_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1.sass.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass,cuda > gpurun_out/$1.src.csv 2>/dev/null
  gzip -f gpurun_out/$1.sass.csv gpurun_out/$1.src.csv
}
prof r02b_cubic_t2_fr4 $PWD/tools/variants/lib_t2m1t32.so cubic 4
prof r02b_linear_lcol_fr2 $PWD/tools/variants/lib_lcolm2t32.so linear 2
ls -la gpurun_out
