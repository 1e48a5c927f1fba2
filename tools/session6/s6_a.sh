set -x
run() { # name, args...
  name=$1; shift
  ncu --set full --clock-control none --import-source on -k regex:jb_process_kernel -c 1 -o /tmp/$name -f python tools/chain_bench.py --steps 1 --warmup 0 --samples 9600 "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.src.csv 2>/dev/null
  gzip -f gpurun_out/$name.src.csv
}
run tex0 --chain JuicyTexture --clips 8192 --synth impulse --path lane --param 0:material=0
run tex1 --chain JuicyTexture --clips 8192 --synth impulse --path lane --param 0:material=1
run tex2 --chain JuicyTexture --clips 8192 --synth impulse --path lane --param 0:material=2
run motion --chain JuicyMotion --clips 16384 --synth drum --path lane
ls -la gpurun_out/
