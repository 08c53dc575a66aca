set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_multifusion.py -m gpu -x -q > gpurun_out/s6_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/s6_tests.log
sweep() {
  python tools/score_bench.py --store --nq 4096 --nv 8197 --k 640 --step 122 --tiles 2,4,1 --reps 9
  python tools/score_bench.py --store --nq 8192 --nv 2056 --k 2048 --step 2432 --tiles 2,4,1 --reps 9
  python tools/score_bench.py --store --nq 8192 --nv 8224 --k 2048 --step 152 --tiles 2,4,1 --reps 9
  python tools/score_bench.py --store --nq 59800 --nv 2990 --k 1536 --tiles 2,4 --reps 5
  python tools/score_bench.py --store --nq 59800 --nv 2990 --k 4608 --tiles 2,4 --reps 5
}
(echo "== NEW"; sweep; python tools/config_bench.py c3 c4) > gpurun_out/s6_store.log 2>&1
cp cross-modal-video-engine_b200/libxmve.so /tmp/new.so; cp tools/_ab/libxmve_before.so cross-modal-video-engine_b200/libxmve.so
(echo "== OLD"; sweep) >> gpurun_out/s6_store.log 2>&1
cp /tmp/new.so cross-modal-video-engine_b200/libxmve.so
grep -v "^+" gpurun_out/s6_store.log | cut -c1-400
