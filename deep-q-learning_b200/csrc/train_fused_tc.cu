// train_fused_tc.cu -- the population form of the fused dueling double-DQN train step with the layer-2 products on the
// tensor cores (sm_100a, tcgen05 + tensor memory).  One CTA of 512 threads = one agent, K steps per launch; the same
// arithmetic as train_fused.cu (General/QLearning/q_agent.py:146-169, q_learning_functions.py:14-64, dddqn.py:24-34).
// Two bodies, chosen per CTA by the agent's batch size (the sweep draws 38..70, hyperparameter_optimization.py:121):
//   train_fused_tc_main.cuh  batches of <= 64 rows (one 64-row tile) and > 80 rows (tiles of 64 rows)
//   train_fused_tc_tail.cuh  batches of 65..80 rows as ONE tile: the up to 16 extra rows ride in lanes and reduction steps
//                            the 64-row tile leaves idle (wider buffers and a wider P1 operand, so it is a body of its own:
//                            the <= 64 case runs 10 % slower through it)
// CTA b runs agent order[b] when the launcher passes an order (costliest agents first: the launch is several waves of
// one-agent CTAs, and the last wave should hold the short ones).
#include "train_fused_tc_main.cuh"
#include "train_fused_tc_tail.cuh"

namespace dqn {

namespace {

constexpr int NT = tcm::NT;
static_assert(tcm::NT == tct::NT, "one block size");

template <int A>
__global__ void __launch_bounds__(NT, 1) dqn_train_tc_kernel(const TrainArgs args) {
  const int sel = args.order ? args.order[blockIdx.x] : (int)blockIdx.x;
  const int B = args.ctl[args.agent_begin + sel].batch_size;       // uniform over the CTA
  if (B > tct::BT && B <= tct::NR) tct::step_body<A>(args, sel);
  else tcm::step_body<A>(args, sel);
}

typedef void (*TrainKernel)(const TrainArgs);
TrainKernel pick_kernel(int A) {
  switch (A) {
    case 2: return dqn_train_tc_kernel<2>;
    case 3: return dqn_train_tc_kernel<3>;
    case 4: return dqn_train_tc_kernel<4>;
    case 5: return dqn_train_tc_kernel<5>;
    case 6: return dqn_train_tc_kernel<6>;
    case 7: return dqn_train_tc_kernel<7>;
    default: return nullptr;
  }
}

}  // namespace

size_t train_tc_smem_bytes(const Dims& d) {
  const size_t a = (size_t)tcm::make_layout(d.D, d.recw).total, b = (size_t)tct::make_layout(d.D, d.recw).total;
  return (a > b ? a : b) * sizeof(float);
}

cudaError_t train_tc_prepare(const Dims& d) {
  TrainKernel k = pick_kernel(d.A);
  if (!k) return cudaErrorInvalidValue;
  return cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_tc_smem_bytes(d));
}

cudaError_t launch_train_tc(cudaStream_t st, const TrainArgs& args) {
  TrainKernel k = pick_kernel(args.dims.A);
  if (!k) return cudaErrorInvalidValue;
  k<<<args.n_sel, NT, train_tc_smem_bytes(args.dims), st>>>(args);
  return cudaGetLastError();
}

}  // namespace dqn
