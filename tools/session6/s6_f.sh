set -x
run() { # name, args...
  name=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:jb_single_kernel -c 1 -o /tmp/$name -f python tools/chain_bench.py --steps 1 --warmup 0 --samples 9600 --path lane "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.src.csv 2>/dev/null
  gzip -f gpurun_out/$name.src.csv
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.summary.json
}
run s_sat --chain JuicySaturator --clips 65536 --synth sweep
run s_infer --chain JuicyInfer --clips 65536 --synth mixed
run s_tex1 --chain JuicyTexture --clips 8192 --synth impulse --param 0:material=1
run s_motion --chain JuicyMotion --clips 16384 --synth drum
run s_satx --chain JuicySaturator --clips 65536 --synth sweep --math exact
ls -la gpurun_out/
