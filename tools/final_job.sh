set -x
mkdir -p gpurun_out
bash tools/gpu_stage.sh tests/test_model_gpu.py tests/test_kernels_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_attention_gpu.py tests/test_edge_cases_gpu.py tests/test_encode_latents_gpu.py tests/test_preprocess_gpu.py tests/test_ddp_gpu.py > gpurun_out/stages.log 2>&1
tail -30 gpurun_out/stages.log | grep -E "stage|passed|failed"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo bench rc=$?; tail -1 gpurun_out/bench_default.log > gpurun_out/r2_bench_d_default.json
timeout 600 python bench.py --config 3 > gpurun_out/bench_c3.log 2>&1; echo bench3 rc=$?; tail -1 gpurun_out/bench_c3.log > gpurun_out/r2_bench_d_train.json
timeout 120 python tools/profile_step.py 64 3 > gpurun_out/ps.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_encode_b64_v2.csv python tools/profile_step.py 64 3 > gpurun_out/ps_ncu.log 2>&1
tail -2 gpurun_out/ps.log
