run() { # name, kernel regex, args...
  name=$1; shift; kre=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:$kre -c 1 -o /tmp/$name -f python tools/chain_bench.py --steps 1 --warmup 0 --samples 9600 --path lane "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.src.csv 2>/dev/null
  gzip -f gpurun_out/$name.src.csv
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.summary.json
}
run p_tex0 jb_pair_kernel --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=0
run t_infer jb_single_kernel --chain JuicyInfer --clips 65536 --synth mixed --inplace
run s_sat2 jb_single_kernel --chain JuicySaturator --clips 65536 --synth sweep --math fast
