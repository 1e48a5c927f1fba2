ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s6_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench_s6.log 2>&1
FULL=JuicyPunch,JuicySaturator,JuicyTexture,JuicyWidth,JuicyMotion,JuicyCohere,JuicyInfer
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_s6_chain.csv python tools/chain_bench.py --steps 1 --warmup 0 --path lane --chain $FULL --clips 32768 --synth mixed > gpurun_out/ncu_chain_s6.log 2>&1
tail -12 gpurun_out/launches_s6_chain.csv | cut -c1-200
