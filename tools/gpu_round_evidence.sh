mkdir -p gpurun_out
T=r02b
(timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -80) > gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${T}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/${T}_launches_bench.log 2>&1
prof() { # name regex target
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/$1 python tools/prof_target.py $3 4 > /dev/null 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
prof ${T}_gemm gemm_bf16 gemm
prof ${T}_reduce reduce_all reduce
prof ${T}_x_gemm_f32_hybrid gemm_tf32 gemm_f32_hybrid
prof ${T}_x_gemm_tf32 gemm_tf32 gemm_f32
prof ${T}_x_gemm_mxf8 gemm_mxf8 gemm_mxf8
prof ${T}_x_gemm_mxf4 gemm_mxf4 gemm_mxf4
prof ${T}_x_argmax reduce_all_argmax argmax
timeout 300 python tools/perf_sweep.py f32 > gpurun_out/${T}_f32_modes.log 2>&1
timeout 300 python tools/perf_sweep.py scaled > gpurun_out/${T}_block_scaled_sweep.log 2>&1
tail -n 6 gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_bench_n1.err | tail -5; head -c 1500 gpurun_out/${T}_bench_n1.json; cat gpurun_out/${T}_f32_modes.log; ls -la gpurun_out | tail -20
