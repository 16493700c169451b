// mlp_kernel2.cuh - the MLP neural-ODE kernel with TWO trajectory tiles in flight per SM (K3 of SURVEY.md §2.1).
//
// Same arithmetic as mlp_kernel.cuh (3xTF32 tcgen05.mma, A from TMEM, W2 resident in shared memory, base-2 activations),
// different schedule.  With one tile per SM the phases of an evaluation are a chain - layer 1 -> the MMAs of the last
// K chunk -> epilogue -> barrier -> RK algebra - so the tensor pipe and the MUFU pipe each idle while the other works
// (measured: ~25 % of the warp time was the mbarrier wait, tensor pipe 35 % active).  Here a CTA carries two independent
// groups, each = one 128-trajectory tile = 8 compute warps + its own MMA warp, named barriers, mbarriers, TMEM region and
// work-queue claims; they share only the read-only weights in shared memory.  While one group waits for its MMAs or a
// barrier, the other group's warps fill the CUDA-core / MUFU pipes and the tensor pipe alternates between the tiles.
//
// TMEM budget: a tile's full A operand (hi + lo, 256 columns) plus D (128) would need 2 x 384 > 512 columns.  A is
// therefore a RING of 4 slots of one K chunk each: compute thread (row, half h) produces its 64 hidden units in 8 chunks
// of 8; chunk c goes to slot c % 4 (hi: 2 halves x 8 columns, lo: the same) = 32 columns per slot; per tile
// D [0,128) + ring [128,256) = 256 columns, two tiles = 512.  The MMA warp commits chunks 0..3 to a per-slot "empty"
// mbarrier, which the producers of chunks 4..7 wait on before overwriting the slot; chunk 7's commit is the "done" mbarrier.
#pragma once
#include "mlp_kernel.cuh"

namespace dfx {

constexpr int kMlp2Groups = 2;
constexpr int kMlp2GroupThreads = 288;  // 8 compute warps + the group's MMA warp
constexpr int kMlp2Threads = kMlp2Groups * kMlp2GroupThreads;
constexpr int kMlp2Chunks = 8, kMlp2Slots = 4;

struct MlpSmem2 {
  float Bhi[kMlpW * kMlpW];     // W2 hi, canonical K-major no-swizzle (see mlp_kernel.cuh)
  float Blo[kMlpW * kMlpW];
  float W1[kMlpW * kMlpD];
  float b1[kMlpW];
  float b2[kMlpW];
  float W3[kMlpW * kMlpD];      // transposed: W3t[o][c]
  float b3[kMlpD];
  float part[kMlp2Groups][2][kMlpW][kMlpD];    // layer-3 partial sums of the two hidden-unit halves
  float k[kMlp2Groups][2][kMlpMaxStages * kMlpD][kMlpW];  // stage values: one private copy per thread of a row (2 halves)
  long long idx[kMlp2Groups][kMlpW];
  unsigned long long mbar_done[kMlp2Groups];
  unsigned long long mbar_empty[kMlp2Groups][kMlp2Slots];
  unsigned long long mbar_w;    // completion of the TMA loads of the W2 image
  uint32_t tmem_base;
};

// named-barrier reductions over one group (PTX barrier.red with an explicit thread count)
__device__ __forceinline__ bool group_or(int id, int count, bool v) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
               : "=r"(r) : "r"((uint32_t)v), "r"(id), "r"(count) : "memory");
  return r != 0;
}
__device__ __forceinline__ bool group_and(int id, int count, bool v) {
  uint32_t r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\tbar.red.and.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
               : "=r"(r) : "r"((uint32_t)v), "r"(id), "r"(count) : "memory");
  return r != 0;
}

template <class Solver, bool FAST_ACT>
__global__ void __launch_bounds__(kMlp2Threads, 1)
mlp_tc2_kernel(const SolveParams<float> p, const float *__restrict__ w, const __grid_constant__ CUtensorMap w2_map) {
  using R = float;
  constexpr int D = kMlpD, W = kMlpW, S = Solver::S;
  static_assert(S <= kMlpMaxStages, "stage-value storage is sized for at most kMlpMaxStages stages");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MlpSmem2 &sm = *reinterpret_cast<MlpSmem2 *>(smem_raw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = warp / 9, lw = warp - 9 * g;  // group, warp within the group (8 = the MMA warp)
  const bool mma_warp = lw == 8;
  // TMEM lane quadrant of a warp is fixed by its rank in the CTA (warp % 4); 4 consecutive warps cover all quadrants
  const int quad = warp & 3, half = (lw >> 2) & 1, row = quad * 32 + lane;
  const int col0 = half * 64;  // this thread's hidden units [col0, col0 + 64)
  // named barriers of this group: 1 + 6 g + {0..3: ring slot full, 4: compute threads, 5: whole group}
  const int bar0 = 1 + 6 * g, bar_cmp = bar0 + 4, bar_grp = bar0 + 5;

  // ---------------- one-time set-up: weights -> smem (W2 split into TF32 hi / lo), TMEM, mbarriers ----------------
  const float *gW1 = w, *gb1 = gW1 + W * D, *gW2 = gb1 + W, *gb2 = gW2 + W * W, *gW3 = gb2 + W, *gb3 = gW3 + D * W;
  // W2 (TF32 hi / lo, already in the UMMA shared-memory layout, mlp_split_w2_kernel): TMA tile loads with mbarrier
  // completion, issued first so the copy engine works while the threads stage the small layers and allocate TMEM
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar_w)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_stage_w2(sm.Bhi, &w2_map, smem_u32(&sm.mbar_w));
  }
  (void)gW2;
  constexpr float kIn = FAST_ACT ? 1.4426950408889634f : 1.0f;   // log2(e) into the pre-activations
  constexpr float kOut = FAST_ACT ? 0.6931471805599453f : 1.0f;  // ln2 back out of the last hidden layer
  for (int i = tid; i < W * D; i += kMlp2Threads) {
    sm.W1[i] = __ldg(gW1 + i) * kIn;
    const int c = i >> 7, o = i & 127;
    sm.W3[o * D + c] = __ldg(gW3 + i) * kOut;
  }
  for (int i = tid; i < W; i += kMlp2Threads) { sm.b1[i] = __ldg(gb1 + i) * kIn; sm.b2[i] = __ldg(gb2 + i) * kIn; }
  if (tid < D) sm.b3[tid] = __ldg(gb3 + tid);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int gg = 0; gg < kMlp2Groups; ++gg) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar_done[gg])) : "memory");
      for (int s = 0; s < kMlp2Slots; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar_empty[gg][s])) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  mbar_wait(smem_u32(&sm.mbar_w), 0);  // the W2 image has landed (TMA complete_tx)
  const uint32_t tmem_all = sm.tmem_base;
  const uint32_t tmem = tmem_all + (uint32_t)(g * 256);          // this tile: D [0,128), A ring [128,256)
  const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);  // this warp's lane quadrant
  const uint32_t mbar_done = smem_u32(&sm.mbar_done[g]);
  const uint64_t bdesc_hi = make_b_desc(smem_u32(sm.Bhi), 128, 4096), bdesc_lo = make_b_desc(smem_u32(sm.Blo), 128, 4096);
  uint32_t phase = 0;  // parity of both the done and the empty barriers: each completes exactly once per evaluation
  float (*gk)[kMlpW] = sm.k[g][half];  // private to this thread: one writer, one reader per word

  // ---------------- the MMA warp's side of one MLP evaluation ----------------
  // chunk c: K-steps kk = 8 h + c (h = 0, 1) of B; A columns of ring slot c % 4: hi at 128 + 32 s + 8 h, lo 16 further
  auto mma_eval = [&]() {
#pragma unroll 1
    for (int c = 0; c < kMlp2Chunks; ++c) {
      const int s = c & (kMlp2Slots - 1);
      named_bar_sync(bar0 + s, kMlp2GroupThreads);  // all 256 producers have stored (and fenced) chunk c
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t bd = (pass == 1) ? bdesc_lo : bdesc_hi;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int kk = 8 * h + c;
            const uint32_t a_col = 128u + 32u * s + 8u * h + ((pass == 2) ? 16u : 0u);
            umma_tf32_ts(tmem, tmem + a_col, bd + (uint64_t)(kk * (256 >> 4)), (c | pass | h) ? 1u : 0u);
          }
        }
        if (c < kMlp2Slots) umma_commit(smem_u32(&sm.mbar_empty[g][s]));  // slot s may be overwritten by chunk c + 4
        if (c == kMlp2Chunks - 1) umma_commit(mbar_done);
      }
      __syncwarp();
    }
  };

  // ---------------- the MLP evaluation, compute threads ----------------
  auto eval = [&](const R (&yin)[D], R (&fout)[D]) {
    auto layer1_chunk = [&](int c, uint32_t (&vh)[8], uint32_t (&vl)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = col0 + c * 8 + j;
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
      uint32_t vh[8], vl[8];
#pragma unroll 1
      for (int c = 0; c < kMlp2Chunks; ++c) {
        const int s = c & (kMlp2Slots - 1);
        layer1_chunk(c, vh, vl);
        if (c > 0) {  // chunk c - 1 is in TMEM: tell the MMA warp (after this chunk's arithmetic, which hid the store latency)
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          named_bar_arrive(bar0 + ((c - 1) & (kMlp2Slots - 1)), kMlp2GroupThreads);
        }
        if (c >= kMlp2Slots) {  // the slot still feeds chunk c - 4's MMAs until their commit arrives
          mbar_wait(smem_u32(&sm.mbar_empty[g][s]), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        tmem_st8(t_lane + 128 + 32 * s + 8 * half, vh);
        tmem_st8(t_lane + 128 + 32 * s + 16 + 8 * half, vl);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      named_bar_arrive(bar0 + ((kMlp2Chunks - 1) & (kMlp2Slots - 1)), kMlp2GroupThreads);
    }
    mbar_wait(mbar_done, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: D row -> + b2 -> softplus -> partial 128 -> 4 output layer over this thread's 64 hidden units
    R acc3[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc3[c] = 0.0f;
#pragma unroll 1
    for (int c32 = 0; c32 < 2; ++c32) {
      uint32_t v0[16], v1[16];
      tmem_ld16(t_lane + col0 + c32 * 32, v0);
      tmem_ld16(t_lane + col0 + c32 * 32 + 16, v1);
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
      out16(col0 + c32 * 32, v0);
      out16(col0 + c32 * 32 + 16, v1);
    }
    *reinterpret_cast<float4 *>(&sm.part[g][half][row][0]) = make_float4(acc3[0], acc3[1], acc3[2], acc3[3]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    named_bar_sync(bar_cmp, kMlp2GroupThreads - 32);
    const float4 p0 = *reinterpret_cast<const float4 *>(&sm.part[g][0][row][0]);
    const float4 p1 = *reinterpret_cast<const float4 *>(&sm.part[g][1][row][0]);
    fout[0] = mlp_tanh<FAST_ACT>((p0.x + p1.x) + sm.b3[0]);
    fout[1] = mlp_tanh<FAST_ACT>((p0.y + p1.y) + sm.b3[1]);
    fout[2] = mlp_tanh<FAST_ACT>((p0.z + p1.z) + sm.b3[2]);
    fout[3] = mlp_tanh<FAST_ACT>((p0.w + p1.w) + sm.b3[3]);
  };

  // ---------------- per-lane trajectory state (identical in the two threads of a row) ----------------
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
    // ---- refill (claims are made by the half-0 thread of each row and shared through smem); group-scoped barriers ----
    if (!exhausted) {
      if (!mma_warp && half == 0) {
        const long long got = claim_work(!active, p.work_counter, (int)(threadIdx.x & 31));
        sm.idx[g][row] = got;
      }
      named_bar_sync(bar_grp, kMlp2GroupThreads);
      bool fail = false;
      if (!mma_warp) {
        const long long got = sm.idx[g][row];
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
      exhausted = group_or(bar_grp, kMlp2GroupThreads, fail);
    }
    if (group_and(bar_grp, kMlp2GroupThreads, mma_warp || !active)) break;

    if (mma_warp) {  // S evaluations per attempted step, group-uniform
#pragma unroll 1
      for (int i = 0; i < S; ++i) { mma_eval(); phase ^= 1u; }
      continue;
    }

    // ---- one attempted step for every lane of the group ----
    const bool run = active && (tprev < t1) && (num_steps < p.max_steps) && (result == DFX_RESULT_SUCCESSFUL);
    const R st0 = tprev, st1 = tnext;
    const R dt = st1 - st0;
    const R control = direction * dt;
    // The stage loop is rolled (one copy of the MLP evaluation in the instruction stream); the stage values therefore
    // live in shared memory, one private copy per thread (the two threads of a row compute identical values), so no
    // barrier is needed and every word has exactly one writer.
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
          for (int c = 0; c < D; ++c) yi[c] += a * gk[j * D + c][row];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) yi[c] = y[c] + yi[c];
      }
      eval(yi, fi);  // the field is autonomous: stage times do not enter
#pragma unroll
      for (int c = 0; c < D; ++c) gk[i * D + c][row] = control * fi[c];
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
        for (int c = 0; c < D; ++c) y1[c] += b * gk[j * D + c][row];
      }
#pragma unroll
      for (int c = 0; c < D; ++c) y1[c] = y[c] + y1[c];
    }
#pragma unroll
    for (int c = 0; c < D; ++c) yerr[c] = 0.f;
    for (int j = 0; j < S; ++j) {
      const R b = Solver::template b_err<R>(j);
#pragma unroll
      for (int c = 0; c < D; ++c) yerr[c] += b * gk[j * D + c][row];
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
      if (half == 0) {
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


  // both groups are done with their TMEM regions before the allocation goes away
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_all), "r"(kTmemCols) : "memory");
}

}  // namespace dfx
