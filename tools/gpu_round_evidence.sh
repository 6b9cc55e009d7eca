mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -80) > gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --quick > gpurun_out/r02_launches_bench.log 2>&1
prof() { # name regex target
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/$1 python tools/prof_target.py $3 4 > /dev/null 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
prof r02_gemm gemm_bf16 gemm
prof r02_reduce reduce_all reduce
prof r02_x_gemm_tf32 gemm_tf32 gemm_f32
prof r02_x_gemm_3xtf32 gemm_tf32 gemm_f32_3x
prof r02_x_reduce_rows reduce_rows reduce_rows
prof r02_x_reduce_cols reduce_cols reduce_cols
prof r02_x_argmax reduce_all_argmax argmax
prof r02_x_gemm_batched gemm_bf16 gemm_batched
tail -n 6 gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_bench_n1.err | tail -5; head -c 1500 gpurun_out/r02_bench_n1.json; ls -la gpurun_out | tail -20
