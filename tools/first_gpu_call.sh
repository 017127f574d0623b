#!/bin/bash
# The first GPU call after a stretch of CPU-only work (run it through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# 1. the whole GPU suite on the default path (the host-side changes made since the last validated build);
# 2. the opt-in negative-binomial path (kernel instantiation + host fold, never run on hardware so far);
# 3. the bench line, default and with the 256-thread instantiation forced (A/B).
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
python tools/sass_hash.py > gpurun_out/sass_hash.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/first_tests.log 2>&1; echo "default suite: exit $?" | tee gpurun_out/first_summary.txt
HFG_EXPERIMENTAL_NB=1 timeout 600 python -m pytest tests/test_gpu_nb_experimental.py -m gpu -q > gpurun_out/first_nb.log 2>&1
echo "negative-binomial (opt-in): exit $?" | tee -a gpurun_out/first_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/first_smoke.log 2>&1; echo "smoke: exit $?" | tee -a gpurun_out/first_summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/first_bench_default.json 2> gpurun_out/first_bench_default.err
HFG_THREADS=256 timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/first_bench_t256.json 2> gpurun_out/first_bench_t256.err
tail -3 gpurun_out/first_tests.log gpurun_out/first_nb.log; cat gpurun_out/first_bench_default.json gpurun_out/first_bench_t256.json
