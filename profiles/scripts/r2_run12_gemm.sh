cd $GRAFT_REPO_ROOT/deep-q-learning_b200/csrc
mkdir -p ../../gpurun_out
for cg in 3 2 1; do echo "== CG=$cg"; DQN_B200_GEMM_CG=$cg timeout 120 ./gemm_tc_test > ../../gpurun_out/r2_gemm_cg$cg.log 2>&1; echo "rc=$?"; grep -E "kind [03]|tcgen05|OK|FAIL|error" ../../gpurun_out/r2_gemm_cg$cg.log | head -14; done
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_large_batch.py -m gpu -q 2>&1 | tail -3
for cg in 3 2 1; do DQN_B200_GEMM_CG=$cg python bench.py --workload dp --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dp cg$cg', d['ms_per_step'], d['roofline']['frac_of_mode_ceiling'])"; done
