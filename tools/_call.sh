mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r1j_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -2 gpurun_out/r1j_attn_tests.log
for st in 0 1000 2000 3000 4000 6000; do echo "stagger $st"; CLM_ATTN_STAGGER=$st timeout 300 python tools/attn_bench.py --iters 200 2>&1; done | tee gpurun_out/attn_bench_r1j.jsonl
