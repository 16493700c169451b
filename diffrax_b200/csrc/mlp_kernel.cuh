// mlp_kernel.cuh - neural-ODE ensemble kernel with the MLP vector field on the 5th-gen tensor cores (K3 of SURVEY.md §2.1).
//
// Field: eqx.nn.MLP(4 -> 128 -> 128 -> 4), softplus hidden, tanh final (docs/examples/neural_ode.ipynb cell 5).
// The 128 x 128 hidden layer is the only real contraction (16 384 of the 17 408 MACs per evaluation); it runs as
//   D[128 traj x 128] = A[128 traj x 128] . W2^T        tcgen05.mma.cta_group::1.kind::tf32, M=128 N=128 K=8
// with fp32-faithful 3xTF32 splitting: x = hi + lo (both rounded to TF32 with cvt.rna), A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi
// accumulated in fp32 in TMEM (residual ~2^-21 relative, the level of fp32 summation noise).
//
// Mapping
//   * one CTA = 128 trajectories = the M tile; 16 compute warps + 1 MMA warp.  Compute thread (warp w, lane l) owns row
//     m = 32 (w % 4) + l (its TMEM lane) and the hidden-unit quarter w / 4 (columns [32 q, 32 q + 32)).  The four threads of
//     a row carry the (tiny, d = 4) RK state redundantly and bit-identically, so only the MLP evaluation communicates.
//   * A (softplus(W1 y + b1), split hi/lo) is written by its owner threads straight into TMEM with tcgen05.st - a thread's
//     row IS its TMEM lane - so A never touches shared memory; W2 (hi and lo, 2 x 64 KB) is resident in shared memory for
//     the whole kernel in the canonical K-major no-swizzle UMMA layout; the accumulator D lives in TMEM and is read back
//     with tcgen05.ld for the epilogue (bias, softplus, the 128 -> 4 output layer, tanh).
//   * The contraction is software-pipelined over K against the layer-1 phase: every thread produces its 32 hidden units
//     in 4 chunks of 8 (= one K = 8 MMA step per quarter); after each chunk it signals a named barrier (bar.arrive, no
//     wait) and the MMA warp (bar.sync on the same barrier) issues that chunk's 4 K-steps x 3 TF32 passes, so the tensor
//     pipe works on chunk c while the CUDA cores compute chunk c + 1; one tcgen05.commit per evaluation releases the
//     epilogue through an mbarrier.
//   * TMEM columns: D [0,128)  A_hi [128,256)  A_lo [256,384)  (512 allocated).
//   * per-trajectory adaptive stepping: every lane has its own t / dt / accept-reject; a finished lane claims the next
//     trajectory from the global queue; the stage loop is CTA-synchronous (one MMA batch per stage evaluation).
//   * every stage, including stage 0, is evaluated each step (the FSAL value f(t1, y1) equals f at the next step's
//     (t0, y0) bit for bit, so re-evaluating it is value-identical to diffrax's reuse; runge_kutta.py:684-695).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#include "ensemble_kernel.cuh"

namespace dfx {

constexpr int kMlpD = 4, kMlpW = 128;
constexpr int kMlpComputeThreads = 512;               // 16 warps: 4 TMEM lane quadrants x 4 hidden-unit quarters
constexpr int kMlpThreads = kMlpComputeThreads + 32;  // + the MMA-issuing warp
constexpr int kMlpChunks = 4, kMlpChunk = 8;          // a thread's 32 hidden units in 4 chunks of 8 (one K = 8 step each)
constexpr uint32_t kTmemCols = 512;

constexpr int kMlpMaxStages = 7;  // tableaux registered for the MLP field: Tsit5, Dopri5 (7), Bosh3 (4), Heun (2)
struct MlpSmem {
  float Bhi[kMlpW * kMlpW];     // W2 hi, canonical K-major no-swizzle: (n,k) -> (n/8)*1024 + (k/4)*32 + (n%8)*4 + (k%4)
  float Blo[kMlpW * kMlpW];
  float W1[kMlpW * kMlpD];
  float b1[kMlpW];
  float b2[kMlpW];
  float W3[kMlpW * kMlpD];     // transposed: W3t[o][c]
  float b3[kMlpD];
  float part[4][kMlpW][kMlpD];  // layer-3 partial sums of the four hidden-unit quarters
  float k[4][kMlpMaxStages * kMlpD][kMlpW];  // stage values k[i][c]: one private copy per thread of a row (4 hidden-unit quarters),
                                             // so every word has exactly one writer and one reader
  long long idx[kMlpW];
  unsigned long long mbar;
  unsigned long long mbar_w;    // completion of the TMA loads of the W2 image
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// UMMA shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor), K-major, SWIZZLE_NONE:
// start address, LBO (K-adjacent core matrices) and SBO (8-row groups) in 16-byte units, version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version_
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// Host: tensor map over the W2 image ([1024 rows][32 floats], row-major, 128-byte rows), box = 256 rows x 32 floats.
// cuTensorMapEncodeTiled is a driver entry point; it is fetched through the runtime so the library links against cudart only.
inline int make_w2_tensor_map(CUtensorMap *out, const float *image_dev, char *err, size_t err_len) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
      snprintf(err, err_len, "cuTensorMapEncodeTiled is not available from this driver (the MLP kernel stages W2 with TMA)");
      return -1;
    }
    encode = (encode_fn)fn;
  }
  const cuuint64_t dims[2] = {32, 2 * kMlpW * kMlpW * 4 / 128};  // innermost first: 32 floats per row, 1024 rows
  const cuuint64_t strides[1] = {128};                            // bytes between rows
  const cuuint32_t box[2] = {32, 256}, estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)image_dev, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, err_len, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -1; }
  return 0;
}

// Ensemble totals (dfx_solve_desc.totals) from the per-trajectory statistics: the tensor-core kernels do not reduce them
// in-kernel, so a small pass over stats / result does (N is 65 536 for BASELINE config 4: microseconds).
__global__ void totals_from_stats_kernel(long long n, const int *__restrict__ stats, const int *__restrict__ result, long long *totals) {
  long long a = 0, b = 0, c = 0, m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int s = stats[3 * i];
    a += s; b += stats[3 * i + 1]; c += (result[i] != DFX_RESULT_SUCCESSFUL && result[i] != DFX_RESULT_EVENT_OCCURRED); m = m > s ? m : s;
  }
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(kFullMask, a, o); b += __shfl_xor_sync(kFullMask, b, o); c += __shfl_xor_sync(kFullMask, c, o);
    const long long mo = __shfl_xor_sync(kFullMask, m, o); m = m > mo ? m : mo;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long *)totals + 0, (unsigned long long)a);
    atomicAdd((unsigned long long *)totals + 1, (unsigned long long)b);
    atomicAdd((unsigned long long *)totals + 2, (unsigned long long)c);
    atomicMax(totals + 3, m);
  }
}

// The fused peer gather for kernels that do not store to the peers themselves (the tensor-core kernels): one pass that copies
// this rank's finals into row peer_row0 + i of every peer buffer (P2P stores over NVLink).
template <class R>
__global__ void peer_scatter_kernel(long long n, int d, const R *__restrict__ y_final, const R *__restrict__ t_final, int n_peers,
                                    long long row0, SolveParams<R> p) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    for (int q = 0; q < n_peers; ++q) {
      for (int c = 0; c < d; ++c) p.peer_y[q][(row0 + i) * d + c] = y_final[i * d + c];
      p.peer_t[q][row0 + i] = t_final[i];
    }
  }
}

// ---- TMA staging of W2 ----
// The 128x128 hidden-layer weights are split into TF32 hi / lo ONCE per launch by a small kernel that writes them to global
// memory already in the shared-memory image the tensor core wants (canonical K-major no-swizzle UMMA layout: 8x4-element
// core matrices of 128 B, LBO 128 B along K, SBO 4096 B along N) - 2 x 64 KB = 1024 rows of 128 B.  Every CTA of the solve
// kernel then stages that image with four cp.async.bulk.tensor (TMA) tile loads of 256 rows each, completion on an
// mbarrier, instead of 32 K per-thread __ldg + st.shared: the copy engine moves the bytes while the threads stage the small
// layers and allocate TMEM.
constexpr int kW2ImageRows = 2 * kMlpW * kMlpW * 4 / 128;  // hi + lo: 1024 rows of 128 bytes
constexpr int kW2BoxRows = 256;                            // one TMA box: 256 rows x 32 floats = 32 KB
__global__ void mlp_split_w2_kernel(const float *__restrict__ gW2, float *__restrict__ image) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kMlpW * kMlpW) return;
  const int n = i >> 7, k = i & 127;  // W2[n][k], (out, in) row-major == K-major B operand
  const float v = __ldg(gW2 + i), hi = to_tf32(v), lo = to_tf32(v - hi);
  const int off = (n >> 3) * 1024 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
  image[off] = hi;
  image[kMlpW * kMlpW + off] = lo;
}
// one 2-D tile load: box (32 floats, kW2BoxRows rows) at row `row0` of the image -> shared memory, completes on `mbar`
__device__ __forceinline__ void tma_load_rows(uint32_t smem_dst, const CUtensorMap *tmap, int row0, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(smem_dst), "l"((unsigned long long)tmap), "r"(0), "r"(row0), "r"(mbar) : "memory");
}
// called by ONE thread after the mbarrier is initialised: arm it with the byte count and issue the four tile loads
__device__ __forceinline__ void tma_stage_w2(float *smem_image, const CUtensorMap *tmap, uint32_t mbar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(kW2ImageRows * 128) : "memory");
#pragma unroll
  for (int r = 0; r < kW2ImageRows; r += kW2BoxRows) tma_load_rows(smem_u32(smem_image) + r * 128, tmap, r, mbar);
}

// UMMA instruction descriptor (InstrDescriptor): D fp32, A/B TF32, both K-major, N=128, M=128.
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kIdescTf32), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}\n" :: "r"(mbar), "r"(parity), "r"(20000u) : "memory");  // suspend-time hint (ns): sleep in hardware instead of spinning on issue slots
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
         "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

// softplus = max(x,0) + log1p(exp(-|x|))   (jax.nn.softplus = logaddexp(x, 0))
// FAST: the kernel works in base 2.  With z' = z log2(e):  softplus(z) = ln2 * sp2(z'),  sp2(z') = max(z',0) + log2(1 + 2^-|z'|)
// (MUFU.EX2 + MUFU.LG2, abs. error ~1e-7).  The log2(e) is folded into W1, b1, b2 and the ln2 into W3 when the weights are
// staged in shared memory; between the two hidden layers the factors cancel (ln2 * log2(e) = 1), so W2 is used as is.
template <bool FAST> __device__ __forceinline__ float mlp_softplus(float x) {
  if constexpr (FAST) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(x)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return fmaxf(x, 0.0f) + l;
  } else {
    return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
  }
}
template <bool FAST> __device__ __forceinline__ float mlp_tanh(float x) {
  if constexpr (FAST) {
    // tanh(x) = 1 - 2 / (1 + e^(2x)); abs. error ~1e-7 (saturates correctly: e -> inf gives 1, e -> 0 gives -1)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return fmaf(-2.0f, r, 1.0f);
  } else {
    return tanhf(x);
  }
}
// TF32 operand split x = hi + lo for 3xTF32.  FAST: hi = x with the 13 low mantissa bits cleared - exactly what the tensor
// core reads of an fp32 word - and lo = x - hi (exact); lo is in turn truncated by the hardware, so the dropped part is
// < 2^-20 |x|.  Otherwise both are rounded to nearest with cvt.rna.tf32 (dropped part < 2^-22 |x|).
template <bool FAST> __device__ __forceinline__ void tf32_split(float x, uint32_t &hi, uint32_t &lo) {
  if constexpr (FAST) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
  } else {
    const float h = to_tf32(x);
    hi = __float_as_uint(h);
    lo = __float_as_uint(to_tf32(x - h));
  }
}

template <class Solver, bool FAST_ACT>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_tc_kernel(const SolveParams<float> p, const float *__restrict__ w, const __grid_constant__ CUtensorMap w2_map) {
  using R = float;
  constexpr int D = kMlpD, W = kMlpW, S = Solver::S;
  static_assert(S <= kMlpMaxStages, "stage-value storage is sized for at most kMlpMaxStages stages");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MlpSmem &sm = *reinterpret_cast<MlpSmem *>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool mma_warp = warp == kMlpComputeThreads / 32;
  const int quad = warp & 3, part = (warp >> 2) & 3, row = quad * 32 + lane;
  const int col0 = part * 32;  // this thread's hidden units [col0, col0 + 32)

  // ---------------- one-time set-up: weights -> smem (W2 split into TF32 hi / lo), TMEM, mbarrier ----------------
  const float *gW1 = w, *gb1 = gW1 + W * D, *gW2 = gb1 + W, *gb2 = gW2 + W * W, *gW3 = gb2 + W, *gb3 = gW3 + D * W;
  // W2 (TF32 hi / lo, already in the UMMA shared-memory layout): TMA tile loads, issued first so they overlap the rest
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar_w)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_stage_w2(sm.Bhi, &w2_map, smem_u32(&sm.mbar_w));
  }
  (void)gW2;
  constexpr float kIn = FAST_ACT ? 1.4426950408889634f : 1.0f;   // log2(e) into the pre-activations
  constexpr float kOut = FAST_ACT ? 0.6931471805599453f : 1.0f;  // ln2 back out of the last hidden layer
  for (int i = tid; i < W * D; i += kMlpThreads) {
    sm.W1[i] = __ldg(gW1 + i) * kIn;
    const int c = i >> 7, o = i & 127;  // W3[c][o] -> W3t[o][c]: one 16-byte load per hidden unit in the epilogue
    sm.W3[o * D + c] = __ldg(gW3 + i) * kOut;
  }
  for (int i = tid; i < W; i += kMlpThreads) { sm.b1[i] = __ldg(gb1 + i) * kIn; sm.b2[i] = __ldg(gb2 + i) * kIn; }
  if (tid < D) sm.b3[tid] = __ldg(gb3 + tid);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  mbar_wait(smem_u32(&sm.mbar_w), 0);  // the W2 image has landed (TMA complete_tx)
  const uint32_t tmem = sm.tmem_base;
  const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);  // this warp's lane quadrant
  const uint32_t mbar = smem_u32(&sm.mbar);
  const uint64_t bdesc_hi = make_b_desc(smem_u32(sm.Bhi), 128, 4096), bdesc_lo = make_b_desc(smem_u32(sm.Blo), 128, 4096);
  uint32_t phase = 0;

  // ---------------- the MMA warp's side of one MLP evaluation ----------------
  // chunk c of every quarter q is K-step kk = 4 q + c; 3xTF32: (A_hi, B_hi), (A_hi, B_lo), (A_lo, B_hi)
  auto mma_eval = [&]() {
#pragma unroll 1
    for (int c = 0; c < kMlpChunks; ++c) {
      named_bar_sync(1 + c, kMlpThreads);  // all 512 producers have stored (and fenced) chunk c of A
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_col = (pass == 2) ? 256u : 128u;
          const uint64_t bd = (pass == 1) ? bdesc_lo : bdesc_hi;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kk = 4 * q + c;
            umma_tf32_ts(tmem, tmem + a_col + kk * 8, bd + (uint64_t)(kk * (256 >> 4)), (c | pass | q) ? 1u : 0u);
          }
        }
        if (c == kMlpChunks - 1) umma_commit(mbar);
      }
      __syncwarp();
    }
  };

  // ---------------- the MLP evaluation, compute threads ----------------
  auto eval = [&](const R (&yin)[D], R (&fout)[D]) {
    // layer 1 (4 -> 128) + softplus on the CUDA cores; split to TF32 hi/lo; straight into TMEM as the A operand.
    // Chunk c is stored, then chunk c + 1 is computed BEFORE waiting for those stores (tcgen05.wait::st) and signalling
    // the MMA warp, so the TMEM store latency hides behind arithmetic.
    auto layer1_chunk = [&](int c, uint32_t (&vh)[kMlpChunk], uint32_t (&vl)[kMlpChunk]) {
#pragma unroll
      for (int j = 0; j < kMlpChunk; ++j) {
        const int o = col0 + c * kMlpChunk + j;
        const float4 w1 = *reinterpret_cast<const float4 *>(&sm.W1[o * D]);
        float acc;
        if constexpr (FAST_ACT) {
          acc = fmaf(w1.x, yin[0], sm.b1[o]);
          acc = fmaf(w1.y, yin[1], acc);
          acc = fmaf(w1.z, yin[2], acc);
          acc = fmaf(w1.w, yin[3], acc);
        } else {
          acc = w1.x * yin[0];
          acc += w1.y * yin[1];
          acc += w1.z * yin[2];
          acc += w1.w * yin[3];
          acc += sm.b1[o];
        }
        tf32_split<FAST_ACT>(mlp_softplus<FAST_ACT>(acc), vh[j], vl[j]);
      }
    };
    {
      uint32_t vh[kMlpChunk], vl[kMlpChunk];
      layer1_chunk(0, vh, vl);
      tmem_st8(t_lane + 128 + col0, vh);
      tmem_st8(t_lane + 256 + col0, vl);
#pragma unroll 1
      for (int c = 1; c < kMlpChunks; ++c) {
        layer1_chunk(c, vh, vl);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        named_bar_arrive(c, kMlpThreads);  // chunk c - 1 is in TMEM (barrier ids 1 .. 4 <-> chunks 0 .. 3)
        tmem_st8(t_lane + 128 + col0 + c * kMlpChunk, vh);
        tmem_st8(t_lane + 256 + col0 + c * kMlpChunk, vl);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      named_bar_arrive(kMlpChunks, kMlpThreads);
    }
    mbar_wait(mbar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: D row -> + b2 -> softplus -> partial 128 -> 4 output layer over this thread's 32 hidden units
    R acc3[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc3[c] = 0.0f;
    {
      uint32_t v0[16], v1[16];
      tmem_ld16(t_lane + col0, v0);  // both halves of this thread's 32 accumulator columns in flight at once
      tmem_ld16(t_lane + col0 + 16, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      auto out16 = [&](int base, const uint32_t (&v)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int o = base + j;
          const float h2 = mlp_softplus<FAST_ACT>(__uint_as_float(v[j]) + sm.b2[o]);
          const float4 w3 = *reinterpret_cast<const float4 *>(&sm.W3[o * D]);
          acc3[0] += w3.x * h2;
          acc3[1] += w3.y * h2;
          acc3[2] += w3.z * h2;
          acc3[3] += w3.w * h2;
        }
      };
      out16(col0, v0);
      out16(col0 + 16, v1);
    }
    *reinterpret_cast<float4 *>(&sm.part[part][row][0]) = make_float4(acc3[0], acc3[1], acc3[2], acc3[3]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar_sync(6, kMlpComputeThreads);
    const float4 p0 = *reinterpret_cast<const float4 *>(&sm.part[0][row][0]);
    const float4 p1 = *reinterpret_cast<const float4 *>(&sm.part[1][row][0]);
    const float4 p2 = *reinterpret_cast<const float4 *>(&sm.part[2][row][0]);
    const float4 p3 = *reinterpret_cast<const float4 *>(&sm.part[3][row][0]);
    fout[0] = mlp_tanh<FAST_ACT>(((p0.x + p1.x) + (p2.x + p3.x)) + sm.b3[0]);
    fout[1] = mlp_tanh<FAST_ACT>(((p0.y + p1.y) + (p2.y + p3.y)) + sm.b3[1]);
    fout[2] = mlp_tanh<FAST_ACT>(((p0.z + p1.z) + (p2.z + p3.z)) + sm.b3[2]);
    fout[3] = mlp_tanh<FAST_ACT>(((p0.w + p1.w) + (p2.w + p3.w)) + sm.b3[3]);
  };

  // ---------------- per-lane trajectory state (identical in the four threads of a row) ----------------
  bool active = false, exhausted = false;
  long long idx = -1;
  R y[D];
  R tprev = 0.f, tnext = 0.f, t0 = 0.f, t1 = 0.f, direction = 1.f, t1_clip_floor = 0.f;
  R pid_inv = 1.f, pid_prev_inv = 1.f;
  bool at_dtmin = false;
  int cs_steps_completed = 1, cs_num_steps = 0;
  int num_steps = 0, num_accepted = 0, result = DFX_RESULT_SUCCESSFUL;
#pragma unroll
  for (int c = 0; c < D; ++c) y[c] = 0.f;
  const R sqrt_d = 2.0f;  // sqrt(D), D == 4

  for (;;) {
    // ---- refill (claims are made by the quarter-0 thread of each row and shared through smem) ----
    if (!exhausted) {
      if (!mma_warp && part == 0) {
        const long long got = claim_work(!active, p.work_counter, (int)(threadIdx.x & 31));
        sm.idx[row] = got;
      }
      __syncthreads();
      bool fail = false;
      if (!mma_warp) {
        const long long got = sm.idx[row];
        fail = !active && got >= p.n_traj;
        if (!active && got >= 0 && got < p.n_traj) {
          idx = got;
          const R a = p.t0_arr ? p.t0_arr[idx] : p.t0, b = p.t1_arr ? p.t1_arr[idx] : p.t1;
          direction = (a < b) ? 1.f : -1.f;
          t0 = a * direction;
          t1 = b * direction;
#pragma unroll
          for (int c = 0; c < D; ++c) y[c] = p.y0[idx * D + c];
          R dt0 = p.has_dt0 ? p.dt0 * direction : 0.01f;  // pid.py:48-49 through WrapTerm
          if (p.controller == DFX_CTRL_PID) {
            if (p.has_dtmax) dt0 = jnp_min(dt0, p.dtmax);
            if (p.has_dtmin) dt0 = jnp_max(dt0, p.dtmin);
          } else {
            const R dt0_up = __int_as_float(__float_as_int(dt0) + (dt0 > 0.f ? 1 : (dt0 < 0.f ? -1 : 1)));
            cs_num_steps = (int)ceil((double)((t1 - t0) / dt0_up));
            cs_steps_completed = 1;
          }
          tprev = t0;
          tnext = jnp_min(t0 + dt0, t1);
          t1_clip_floor = prev_n<R>(t1, 100);
          pid_inv = 1.f; pid_prev_inv = 1.f; at_dtmin = false;
          num_steps = 0; num_accepted = 0; result = DFX_RESULT_SUCCESSFUL;
          active = true;
        }
      }
      exhausted = __syncthreads_or(fail) != 0;
    }
    if (__syncthreads_and(mma_warp || !active)) break;

    if (mma_warp) {  // S evaluations per attempted step, CTA-uniform
#pragma unroll 1
      for (int i = 0; i < S; ++i) mma_eval();
      continue;
    }

    // ---- one attempted step for every lane of the CTA ----
    const bool run = active && (tprev < t1) && (num_steps < p.max_steps) && (result == DFX_RESULT_SUCCESSFUL);
    const R st0 = tprev, st1 = tnext;
    const R dt = st1 - st0;
    const R control = direction * dt;
    // The stage loop is rolled (one copy of the MLP evaluation in the instruction stream); the stage values therefore
    // live in shared memory, one copy per row: the four threads of a row store bit-identical values to the same word and
    // each reads back what it wrote itself, so no barrier is needed.
    R y1[D], yerr[D], yi[D], fi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) yi[c] = y[c];
#pragma unroll 1
    for (int i = 0; i < S; ++i) {
      if (i > 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) yi[c] = 0.f;
        for (int j = 0; j < i; ++j) {
          const R a = Solver::template a<R>(i, j);  // structural zeros contribute exact zeros
#pragma unroll
          for (int c = 0; c < D; ++c) yi[c] += a * sm.k[part][j * D + c][row];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) yi[c] = y[c] + yi[c];
      }
      eval(yi, fi);  // the field is autonomous: stage times do not enter
#pragma unroll
      for (int c = 0; c < D; ++c) sm.k[part][i * D + c][row] = control * fi[c];
    }
    if constexpr (Solver::kSsal) {
#pragma unroll
      for (int c = 0; c < D; ++c) y1[c] = yi[c];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) y1[c] = 0.f;
      for (int j = 0; j < S; ++j) {
        const R b = Solver::template b_sol<R>(j);
#pragma unroll
        for (int c = 0; c < D; ++c) y1[c] += b * sm.k[part][j * D + c][row];
      }
#pragma unroll
      for (int c = 0; c < D; ++c) y1[c] = y[c] + y1[c];
    }
#pragma unroll
    for (int c = 0; c < D; ++c) yerr[c] = 0.f;
    for (int j = 0; j < S; ++j) {
      const R b = Solver::template b_err<R>(j);
#pragma unroll
      for (int c = 0; c < D; ++c) yerr[c] += b * sm.k[part][j * D + c][row];
    }

    if (run) {
      bool keep;
      R next_t0, next_t1;
      if (p.controller == DFX_CTRL_PID) {  // pid.py:394-567 (faithful fp32 path)
        bool nan_any = false;
#pragma unroll
        for (int c = 0; c < D; ++c) nan_any |= r_isnan(y1[c]);
        R ss = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const R e = r_isnan(yerr[c]) ? Num<R>::inf() : yerr[c];
          const R yc = nan_any ? y[c] : y1[c];
          const R sc = e / (p.atol + fmaxf(fabsf(y[c]), fabsf(yc)) * p.rtol);
          ss += sc * sc;
        }
        const R scaled_error = sqrtf(ss) / sqrt_d;
        keep = scaled_error < 1.f;
        if (p.has_dtmin) keep = keep || at_dtmin;
        R inv = 1.f / scaled_error;
        R factor = p.safety;
        if (p.use_c1) factor = factor * powf(inv, p.coeff1);
        if (p.use_c2) factor = factor * powf(pid_inv, p.coeff2);
        if (p.use_c3) factor = factor * powf(pid_prev_inv, p.coeff3);
        factor = jnp_min(jnp_max(factor, keep ? 1.f : p.factormin), keep ? p.factormax : p.safety);
        R dtn = dt * factor;
        if (inv == 0.f || r_isinf(inv)) inv = 1.f;
        if (p.has_dtmax) dtn = jnp_min(dtn, p.dtmax);
        if (p.has_dtmin) {
          if (!p.force_dtmin && dtn < p.dtmin && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_DT_MIN_REACHED;
          if (at_dtmin && factor == 1.f) dtn = p.dtmin;
          at_dtmin = dtn <= p.dtmin;
          dtn = jnp_max(dtn, p.dtmin);
        }
        next_t0 = keep ? st1 : st0;
        next_t1 = next_t0 + dtn;
        if (keep) { pid_prev_inv = pid_inv; pid_inv = inv; }
      } else {  // constant.py:57-104
        keep = true;
        cs_steps_completed += 1;
        R t1n = t0 + (t1 - t0) * ((R)cs_steps_completed / (R)cs_num_steps);
        if (cs_steps_completed == cs_num_steps) t1n = t1;
        next_t0 = st1;
        next_t1 = t1n;
      }
      const R tprev_new = next_t0;
      R tnext_new = next_t1;
      if (next_t1 > t1_clip_floor) tnext_new = keep ? t1 : tprev_new + 0.5f * (t1 - tprev_new);
      num_steps += 1;
      num_accepted += keep ? 1 : 0;
#pragma unroll
      for (int c = 0; c < D; ++c) y[c] = keep ? y1[c] : y[c];
      tprev = tprev_new;
      tnext = tnext_new;
    }
    const bool finished = active && !((tprev < t1) && (num_steps < p.max_steps) && (result == DFX_RESULT_SUCCESSFUL));
    if (finished) {
      if ((tprev < t1) && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_MAX_STEPS_REACHED;
      if (part == 0) {
        if (p.save_t1) {
          p.ts_out[idx] = tprev * direction;
          *reinterpret_cast<float4 *>(&p.ys_out[idx * D]) = make_float4(y[0], y[1], y[2], y[3]);
        }
        p.stats[idx * 3 + 0] = num_steps;
        p.stats[idx * 3 + 1] = num_accepted;
        p.stats[idx * 3 + 2] = num_steps - num_accepted;
        p.result[idx] = result;
        if (p.y_final) *reinterpret_cast<float4 *>(&p.y_final[idx * D]) = make_float4(y[0], y[1], y[2], y[3]);
        if (p.t_final) p.t_final[idx] = tprev * direction;
      }
      active = false;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace dfx
