// inst_mlp.cu - neural-ODE MLP field (d=4, width 128, depth 2, fp32).
//   SaveAt(t1=True) solves run on the tensor-core kernel (mlp_kernel.cuh: tcgen05.mma kind::tf32, 3xTF32, A and D in TMEM);
//   every other SaveAt mode (and DFX_MLP_NO_TC=1) runs the exact-fp32 CUDA-core functor inside the generic kernel.
#include <cstdlib>

#include "launch.cuh"
#include "mlp_kernel.cuh"
#include "mlp_kernel2.cuh"

namespace {
using namespace dfx;
using F = MlpField<4, 128>;

template <class Solver>
int launch_mlp(const dfx_solve_desc *d, void *stream_v) {
  const bool rich = d->save_t0 || d->save_ts || d->save_steps || d->save_dense || d->step_ts || d->jump_ts ||
                    d->hairer_initial_step || d->n_events != 0 || d->state_in || d->state_out || d->store_rejected_steps > 0;
  const char *no_tc = std::getenv("DFX_MLP_NO_TC");
  if (rich || !IsTableau<Solver>::value || (no_tc && no_tc[0] == '1')) return launch_solve<float, F, Solver, 0>(d, stream_v);
  if constexpr (IsTableau<Solver>::value) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    if ((int)d->field_params[0] != 128 || d->n_field_weights != F::kNumWeights) {
      set_error("MLP tensor-core kernel: expected width 128 / %lld weights, got width %d / %lld weights", F::kNumWeights,
                (int)d->field_params[0], (long long)d->n_field_weights);
      return DFX_ERR_BAD_ARGUMENT;
    }
    SolveParams<float> p;
    fill_params<float, Solver>(d, p, false);
    if (p.n_traj == 0) return 0;
    unsigned long long *counter = nullptr;
    DFX_CUDA_OK(cudaMallocAsync((void **)&counter, 16, stream));
    DFX_CUDA_OK(cudaMemsetAsync(counter, 0, 16, stream));
    p.work_counter = counter;
    // W2 -> TF32 hi / lo in the UMMA shared-memory image (global scratch), and the tensor map the CTAs stage it through (TMA)
    float *w2_image = nullptr;
    DFX_CUDA_OK(cudaMallocAsync((void **)&w2_image, 2 * sizeof(float) * kMlpW * kMlpW, stream));
    const float *gW2 = (const float *)d->field_weights + kMlpW * kMlpD + kMlpW;
    mlp_split_w2_kernel<<<(kMlpW * kMlpW + 255) / 256, 256, 0, stream>>>(gW2, w2_image);
    count_launch();
    CUtensorMap w2_map;
    char tm_err[160];
    if (make_w2_tensor_map(&w2_map, w2_image, tm_err, sizeof tm_err)) {
      set_error("%s", tm_err);
      cudaFreeAsync(w2_image, stream);
      cudaFreeAsync(counter, stream);
      return DFX_ERR_CUDA;
    }
    const char *slow_act = std::getenv("DFX_MLP_EXACT_ACT");
    const bool fast = !(slow_act && slow_act[0] == '1');
    int sms = 0;
    if (int rc = device_sm_count(&sms)) return rc;
    const char *tiles_env = std::getenv("DFX_MLP_TILES");  // 2 (default): two tiles in flight per SM, mlp_kernel2.cuh; 1: mlp_kernel.cuh
    if (tiles_env && tiles_env[0] == '1') {
      auto kern = fast ? mlp_tc_kernel<Solver, true> : mlp_tc_kernel<Solver, false>;
      const int smem = (int)sizeof(MlpSmem);
      DFX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      long long blocks = (p.n_traj + 127) / 128;
      if (blocks > sms) blocks = sms;  // persistent: one CTA (one 128-trajectory M tile) per SM
      kern<<<(unsigned)blocks, kMlpThreads, smem, stream>>>(p, (const float *)d->field_weights, w2_map);
    } else {
      auto kern = fast ? mlp_tc2_kernel<Solver, true> : mlp_tc2_kernel<Solver, false>;
      const int smem = (int)sizeof(MlpSmem2);
      DFX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      long long blocks = (p.n_traj + 255) / 256;
      if (blocks > sms) blocks = sms;  // persistent: one CTA (two 128-trajectory M tiles) per SM
      kern<<<(unsigned)blocks, kMlp2Threads, smem, stream>>>(p, (const float *)d->field_weights, w2_map);
    }
    count_launch();
    DFX_CUDA_OK(cudaGetLastError());
    if (p.n_peers != 0) {
      if (!p.y_final || !p.t_final) { set_error("the MLP kernel's peer gather needs y_final / t_final buffers"); return DFX_ERR_BAD_ARGUMENT; }
      peer_scatter_kernel<float><<<148, 256, 0, stream>>>(p.n_traj, kMlpD, p.y_final, p.t_final, p.n_peers, p.peer_row0, p);
      count_launch();
    }
    if (p.totals) {
      DFX_CUDA_OK(cudaMemsetAsync(p.totals, 0, 4 * sizeof(long long), stream));
      totals_from_stats_kernel<<<64, 256, 0, stream>>>(p.n_traj, p.stats, p.result, p.totals);
      count_launch();
    }
    cudaFreeAsync(counter, stream);
    cudaFreeAsync(w2_image, stream);
  }
  return 0;
}

template <class Solver>
struct MlpRegistrar {
  MlpRegistrar() { register_builtin(F::kId, F::kDim, Solver::kId, DFX_F32, 0, &launch_mlp<Solver>); }
};
static MlpRegistrar<Tsit5> r0;
static MlpRegistrar<Dopri5> r1;
static MlpRegistrar<Bosh3> r2;
static MlpRegistrar<Heun> r3;
static MlpRegistrar<EulerSolver> r4;
}  // namespace
