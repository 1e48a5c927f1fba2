run() { # name, args...
  name=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:jb_single_kernel -c 1 -o /tmp/$name -f python tools/chain_bench.py --steps 1 --warmup 0 --samples 9600 --path lane "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.src.csv 2>/dev/null
  gzip -f gpurun_out/$name.src.csv
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.summary.json
}
run s_infer2 --chain JuicyInfer --clips 65536 --synth mixed
