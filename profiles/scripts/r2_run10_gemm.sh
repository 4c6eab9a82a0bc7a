cd $GRAFT_REPO_ROOT/deep-q-learning_b200/csrc
mkdir -p ../../gpurun_out
echo "== CG=2"; DQN_B200_GEMM_DEBUG=1 timeout 120 ./gemm_tc_test > ../../gpurun_out/r2_gemm_cg2.log 2>&1; echo "rc=$?"; cat ../../gpurun_out/r2_gemm_cg2.log
echo "== CG=1"; DQN_B200_GEMM_CG=1 timeout 120 ./gemm_tc_test > ../../gpurun_out/r2_gemm_cg1.log 2>&1; echo "rc=$?"; grep -A2 "16384" ../../gpurun_out/r2_gemm_cg1.log
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
