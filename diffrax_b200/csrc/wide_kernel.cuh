// wide_kernel.cuh - one trajectory per WARP, for wide states (north star: "one trajectory per thread, or per warp for wide
// states").  The one-thread-per-trajectory kernel keeps y and the stage values k[S][d] of a trajectory in ONE thread's
// registers, which stops at d = 8; here the state is spread over the 32 lanes of a warp - lane l owns the components
// l, l + 32, l + 64, ... - so a trajectory has 32 x the registers (d up to 1024) and the explicit RK algebra, which is
// componentwise (runge_kutta.py:843-1069: y_i = y0 + sum_j a_ij k_j, y_error = sum_j b_err_j k_j; the local interpolants),
// runs on all lanes in parallel without communication.  Two things couple the components:
//   * the vector field: the warp publishes the stage state in shared memory (d values), then every lane evaluates ITS
//     components of f(t, y) from the full vector - Field::component<R>(params, t, i, y);
//   * the norms of the step-size controller (optx.rms_norm, pid.py:492): per-lane partial sums + a butterfly reduction, which
//     leaves the same value in every lane, so that all control flow (accept / reject, the step loop, the work queue) is
//     warp-uniform: no divergence, and the trajectories of different warps are as independent as in the other kernel.
// Scope: explicit RK tableaux (Tsit5, Dopri5, Dopri8, Bosh3, Heun, Midpoint, Ralston), PIDController (the faithful
// pid.py:394-567 path; dtmin / dtmax) and ConstantStepSize, SaveAt(t0, t1, ts, steps), per-trajectory t0 / t1, finals + totals + the
// fused peer gather of the sharded entry.
// The controller is the reference's faithful path; the stage sums are chained onto y0 for <= 7 stages like in the per-thread
// kernel (DFX_OPT_CHAIN_Y0=0 restores vector_tree_dot, then y0 + incr).
#pragma once
#include "launch.cuh"

namespace dfx {

template <class R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // (x + y == y + x: every lane ends with the same bits)
  return v;
}

constexpr int kWideBlock = 128;
// Occupancy target for ptxas.  Measured (Lorenz-96 / Dopri5 / fp64, 32 768 trajectories): D = 64: 4.90 ms unconstrained (156
// registers, 3 CTAs/SM), 4.20 at 4 CTAs/SM, 4.08 at 5; D = 128: 6.79 / 6.13 / 6.26; D = 256 (8 components per lane): 12.9 /
// 14.9 / 22.9 (spills) - so: 4 CTAs/SM up to 4 components per lane, unconstrained above.  DFX_WIDE_MIN_BLOCKS overrides.
#ifndef DFX_WIDE_MIN_BLOCKS
#define DFX_WIDE_MIN_BLOCKS 0
#endif
template <int D> constexpr int wide_min_blocks() { return DFX_WIDE_MIN_BLOCKS > 0 ? DFX_WIDE_MIN_BLOCKS : (D <= 128 ? 4 : 1); }

template <class R, class Field, class Solver>
__global__ void __launch_bounds__(kWideBlock, wide_min_blocks<Field::kDim>()) wide_kernel(SolveParams<R> p, typename Field::template P<R> fp_in) {
  [[maybe_unused]] typename Field::template P<R> fp_warp = fp_in;   // per-trajectory args (see ensemble_kernel.cuh)
  const typename Field::template P<R> &fp = PerTrajArgs<Field>::value ? fp_warp : fp_in;
  constexpr int D = Field::kDim, S = Solver::S, CH = (D + 31) / 32;
  constexpr bool FSAL = Solver::kFsal;
  constexpr int INTERP = Solver::kInterp;
  extern __shared__ __align__(16) unsigned char wide_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  R *ysh = reinterpret_cast<R *>(wide_smem_raw) + (size_t)warp * D;
  const R sqrt_d = (R)sqrt((double)D);

  // f(t, y) for this lane's components; y_lane holds them (slots past the end of the state stay zero everywhere)
  auto feval = [&](R t, const R (&yy)[CH], R (&f)[CH]) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (lane + 32 * j < D) ysh[lane + 32 * j] = yy[j];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if ((D % 32 == 0) || lane + 32 * j < D) f[j] = Field::template component<R>(fp, t, lane + 32 * j, ysh);
      else f[j] = R(0);
    }
  };
  auto save_row = [&](long long idx, int slot, R t_user, const R (&yy)[CH]) {
    const long long o = idx * (long long)p.out_size + slot;
    if (lane == 0) p.ts_out[o] = t_user;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (lane + 32 * j < D) p.ys_out[o * D + lane + 32 * j] = yy[j];
  };

  long long tot_steps = 0, tot_acc = 0, tot_fail = 0, tot_max = 0;  // this warp's share of p.totals (identical in every lane)
  for (;;) {
    unsigned long long got = 0;
    if (lane == 0) got = atomicAdd(p.work_counter, 1ull);
    got = __shfl_sync(0xffffffffu, got, 0);
    if ((long long)got >= p.n_traj) break;
    const long long idx = (long long)got;

    // ---- set-up: _integrate.py:1076-1079, 1157-1165 (direction), pid.py:316-392 / constant.py:30-55 (first step) ----
    const R a_ = p.t0_arr ? p.t0_arr[idx] : p.t0, b_ = p.t1_arr ? p.t1_arr[idx] : p.t1;
    const R direction = (a_ < b_) ? R(1) : R(-1);
    const R t0 = a_ * direction, t1 = b_ * direction;
    R y[CH], f_fsal[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) { y[j] = (lane + 32 * j < D) ? p.y0[idx * D + lane + 32 * j] : R(0); f_fsal[j] = R(0); }
    if constexpr (PerTrajArgs<Field>::value) {
      if (p.traj_args != nullptr) {
#pragma unroll
        for (int i = 0; i < Field::kNumParams; ++i) fp_warp.p[i] = p.traj_args[idx * Field::kNumParams + i];
      }
    }
    R dt0 = p.has_dt0 ? p.dt0 * direction : R(0.01);
    int cs_num_steps = 0;
    if (p.controller == DFX_CTRL_PID) {
      if (p.has_dtmax) dt0 = jnp_min(dt0, p.dtmax);
      if (p.has_dtmin) dt0 = jnp_max(dt0, p.dtmin);
    } else {
      const R dt0_up = Num<R>::from_bits(Num<R>::bits(dt0) + (dt0 > R(0) ? 1 : (dt0 < R(0) ? -1 : 1)));
      cs_num_steps = r_isinf(t1) ? -1 : (int)ceil((double)((t1 - t0) / dt0_up));
    }
    R tprev = t0, tnext = jnp_min(t0 + dt0, t1);                  // _integrate.py:1265
    const R t1_clip_floor = prev_n<R>(t1, 100);                    // _integrate.py:320-322
    R pid_inv = R(1), pid_prev_inv = R(1);
    bool at_dtmin = false;
    int num_steps = 0, num_accepted = 0, result = DFX_RESULT_SUCCESSFUL, save_index = 0, saveat_ts_index = 0;
    if constexpr (FSAL) feval(t0 * direction, y, f_fsal);          // runge_kutta.py:684-695 (first_step)
    if (p.save_t0) { save_row(idx, 0, t0 * direction, y); save_index = 1; }   // _integrate.py:329-341

    // ---- the step loop, _integrate.py:355-363, 685-687 ----
    while ((tprev < t1) && (num_steps < p.max_steps) && result == DFX_RESULT_SUCCESSFUL) {
      const R st0 = tprev, st1 = tnext, dt = st1 - st0;
      const R control = direction * dt;                            // WrapTerm.contr (_term.py:742-745)
      R k[S][CH], yi[CH], fi[CH], y1[CH], yerr[CH];
      // ---- explicit RK step, runge_kutta.py:643-1203 ----
      if constexpr (FSAL) {
#pragma unroll
        for (int j = 0; j < CH; ++j) k[0][j] = control * f_fsal[j];
      } else {
        feval(st0 * direction, y, fi);
#pragma unroll
        for (int j = 0; j < CH; ++j) k[0][j] = control * fi[j];
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) { yi[j] = y[j]; fi[j] = f_fsal[j]; }
#pragma unroll
      for (int i = 1; i < S; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          if constexpr (DFX_OPT_CHAIN_Y0 && S <= 7) {
            // as in the per-thread kernel: ONE chain of FMAs seeded with y0 (a DMUL and a DADD less per stage and component)
            R acc = y[c];
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (Solver::hA(i * (i - 1) / 2 + j) != 0.0) acc += Solver::template a<R>(i, j) * k[j][c];
            yi[c] = acc;
          } else {
            R incr = R(0);
#pragma unroll
            for (int j = 0; j < i; ++j)
              if (Solver::hA(i * (i - 1) / 2 + j) != 0.0) incr += Solver::template a<R>(i, j) * k[j][c];  // vector_tree_dot (base.py:37-41)
            yi[c] = y[c] + incr;                                   // 871
          }
        }
        const R ti = (Solver::hC(i) == 1.0) ? st1 : st0 + Solver::template c<R>(i) * dt;  // 1023
        feval(ti * direction, yi, fi);
#pragma unroll
        for (int c = 0; c < CH; ++c) k[i][c] = control * fi[c];
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if constexpr (Solver::kSsal) y1[c] = yi[c];                // 1161
        else {                                                     // 1177-1185
          R incr = R(0);
#pragma unroll
          for (int j = 0; j < S; ++j)
            if (Solver::hBsol(j) != 0.0) incr += Solver::template b_sol<R>(j) * k[j][c];
          y1[c] = y[c] + incr;
        }
        R e = R(0);                                                // 1186-1193
#pragma unroll
        for (int j = 0; j < S; ++j)
          if (Solver::hBerr(j) != 0.0) e += Solver::template b_err<R>(j) * k[j][c];
        yerr[c] = e;
      }

      // ---- step-size controller ----
      bool keep;
      R next_t0, next_t1;
      if (p.controller == DFX_CTRL_PID) {                          // pid.py:394-567; y_error NaN -> inf first (_integrate.py:386)
        bool nan_lane = false;
#pragma unroll
        for (int c = 0; c < CH; ++c) nan_lane |= r_isnan(y1[c]);
        const bool nan_any = __any_sync(0xffffffffu, nan_lane);
        R ss = R(0);
#pragma unroll
        for (int c = 0; c < CH; ++c) {                             // _scale, 483-490
          const R e = r_isnan(yerr[c]) ? Num<R>::inf() : yerr[c];
          const R yc = nan_any ? y[c] : y1[c];
          const R yy = r_max(r_abs(y[c]), r_abs(yc));
          const R sc = e / (p.atol + yy * p.rtol);
          ss += sc * sc;
        }
        ss = warp_sum(ss);
        const R scaled_error = (D == 1) ? r_sqrt(ss) : r_sqrt(ss) / sqrt_d;  // optx.rms_norm
        keep = scaled_error < R(1);                                // 493
        if (p.has_dtmin) keep = keep || at_dtmin;                  // 495-496
        R inv = R(1) / scaled_error;                               // 498
        R factor = p.safety;
        if (p.use_c1) factor = factor * r_pow(inv, p.coeff1);      // 515
        if (p.use_c2) factor = factor * r_pow(pid_inv, p.coeff2);  // 516
        if (p.use_c3) factor = factor * r_pow(pid_prev_inv, p.coeff3);  // 517
        const R fmin = keep ? R(1) : p.factormin, fmax = keep ? p.factormax : p.safety;  // 518-520
        factor = jnp_min(jnp_max(factor, fmin), fmax);             // 521-525
        R dtn = x_mul(dt, factor);                                 // 531
        if (inv == R(0) || r_isinf(inv)) inv = R(1);               // 537-538
        if (p.has_dtmax) dtn = jnp_min(dtn, p.dtmax);              // 545-546
        if (p.has_dtmin) {                                         // 547-555
          if (!p.force_dtmin && dtn < p.dtmin && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_DT_MIN_REACHED;
          if (at_dtmin && factor == R(1)) dtn = p.dtmin;
          at_dtmin = dtn <= p.dtmin;
          dtn = jnp_max(dtn, p.dtmin);
        }
        next_t0 = keep ? st1 : st0;                                // 557-558
        next_t1 = x_add(next_t0, dtn);
        if (keep) { pid_prev_inv = pid_inv; pid_inv = inv; }       // 560-564
      } else {                                                     // constant.py:57-104
        keep = true;
        const int done = num_steps + 2;
        R t1n = t0 + (t1 - t0) * ((R)done / (R)cs_num_steps);
        if (done == cs_num_steps) t1n = t1;
        if (cs_num_steps < 0) t1n = st1 + p.dt0 * direction;
        next_t0 = st1;
        next_t1 = t1n;
      }
      // ---- book-keeping, _integrate.py:412-437, 278-284 ----
      const R tprev_new = next_t0;
      R tnext_new = next_t1;
      if (next_t1 > t1_clip_floor) tnext_new = keep ? t1 : tprev_new + R(0.5) * (t1 - tprev_new);
      num_steps += 1;
      num_accepted += keep ? 1 : 0;

      // ---- SaveAt(ts): the step's interpolant, kept steps only (456-487); SaveAt(steps=n) (493-524) ----
      if (p.save_ts != nullptr && keep) {
        while (saveat_ts_index < p.n_save_ts) {
          const R tq = p.save_ts[saveat_ts_index] * direction;
          if (!(tq <= st1)) break;
          R yq[CH];
          interp_eval<INTERP, R, S, CH>(st0, st1, y, y1, k, tq, yq);
          save_row(idx, save_index, tq * direction, yq);
          saveat_ts_index += 1;
          save_index += 1;
        }
      }
      if (p.save_steps != 0 && keep && (num_accepted % p.save_steps) == 0) {
        save_row(idx, save_index, tprev_new * direction, y1);
        save_index += 1;
      }
      if (keep) {  // (warp-uniform: a branch, not 2 CH selects)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          y[c] = y1[c];
          if constexpr (FSAL) f_fsal[c] = fi[c];                   // 1199-1200: the last stage's derivative is f(t1, y1)
        }
      }
      tprev = tprev_new;
      tnext = tnext_new;
    }

    // ---- finalise: _integrate.py:823-883 ----
    if (t0 == t1 && p.save_ts != nullptr) {
      for (int i = 0; i < p.n_save_ts; ++i) { save_row(idx, save_index, t0 * direction, y); save_index += 1; }
    }
    {
      bool via_steps = false;
      if (p.save_steps == 1) via_steps = true;
      else if (p.save_steps > 1) via_steps = (num_accepted % p.save_steps) == 0;
      if (p.save_t1 && !via_steps && save_index < p.out_size) { save_row(idx, save_index, tprev * direction, y); save_index += 1; }
    }
    if ((tprev < t1) && result == DFX_RESULT_SUCCESSFUL) result = DFX_RESULT_MAX_STEPS_REACHED;
    if (lane == 0) {
      p.stats[idx * 3 + 0] = num_steps; p.stats[idx * 3 + 1] = num_accepted; p.stats[idx * 3 + 2] = num_steps - num_accepted;
      p.result[idx] = result;
      if (p.save_count) p.save_count[idx] = save_index;
      if (p.t_final) p.t_final[idx] = tprev * direction;
    }
    if (p.y_final) {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (lane + 32 * j < D) p.y_final[idx * D + lane + 32 * j] = y[j];
    }
    if (p.n_peers != 0) {
      // the all_gather of the finals, fused (see ensemble_kernel.cuh): P2P stores into every rank's global buffer
      const long long g = p.peer_row0 + idx;
#pragma unroll 1
      for (int q = 0; q < p.n_peers; ++q) {
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (lane + 32 * j < D) p.peer_y[q][g * D + lane + 32 * j] = y[j];
        if (lane == 0) p.peer_t[q][g] = tprev * direction;
      }
    }
    // unfilled output slots read +inf (_integrate.py:1296-1300, 1320-1322)
    if (p.save_ts != nullptr || p.save_steps > 0) {
      pad_tail(p.ts_out + idx * p.out_size, (long long)save_index, (long long)p.out_size, lane, false);
      pad_tail(p.ys_out + idx * p.out_size * D, (long long)save_index * D, (long long)p.out_size * D, lane, false);
    }
    tot_steps += num_steps; tot_acc += num_accepted;
    tot_fail += (result != DFX_RESULT_SUCCESSFUL && result != DFX_RESULT_EVENT_OCCURRED) ? 1 : 0;
    tot_max = tot_max > num_steps ? tot_max : (long long)num_steps;
  }
  if (p.totals != nullptr && lane == 0 && (tot_steps | tot_fail) != 0) {
    atomicAdd((unsigned long long *)&p.totals[0], (unsigned long long)tot_steps);
    atomicAdd((unsigned long long *)&p.totals[1], (unsigned long long)tot_acc);
    atomicAdd((unsigned long long *)&p.totals[2], (unsigned long long)tot_fail);
    atomicMax((long long *)&p.totals[3], tot_max);
  }
}

// The launcher registered for a wide functor: same descriptor, same checks, one warp per trajectory.
template <class R, class Field, class Solver>
int launch_wide(const dfx_solve_desc *d, void *stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  static_assert(IsTableau<Solver>::value && !IsHalf<Solver>::value, "the wide kernel runs explicit RK tableaux");
  if (d->levy_area != DFX_LEVY_NONE || d->n_events != 0 || d->step_ts || d->jump_ts || d->state_in || d->state_out || d->save_dense ||
      d->store_rejected_steps > 0 || (d->hairer_initial_step && std::isnan(d->dt0))) {
    set_error("the warp-per-trajectory kernel of a wide functor covers ODE solves with SaveAt(t0, t1, ts, steps): no Brownian "
              "motion, events, ClipStepSizeController, dense output, resumed states or Hairer starting step");
    return DFX_ERR_UNSUPPORTED;
  }
  if (host_pipe() != nullptr) { set_error("internal: the host pipeline does not drive the wide kernel"); return DFX_ERR_UNSUPPORTED; }
  SolveParams<R> p;
  fill_params<R, Solver>(d, p, false);
  if (d->n_field_params < Field::kNumParams) { set_error("field needs %d parameters, got %d", Field::kNumParams, d->n_field_params); return DFX_ERR_BAD_ARGUMENT; }
  const auto fp = Field::template make<R>(d->field_params, d->n_field_params, d->field_weights);
  if (p.n_traj == 0) return 0;
  if (d->traj_args != nullptr && (!PerTrajArgs<Field>::value || d->n_traj_args != Field::kNumParams)) {
    set_error("per-trajectory args: this functor takes %d (got %d)", PerTrajArgs<Field>::value ? Field::kNumParams : 0, d->n_traj_args);
    return DFX_ERR_BAD_ARGUMENT;
  }
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  const size_t smem = (size_t)(kWideBlock / 32) * Field::kDim * sizeof(R);
  int per_sm = 0;
  DFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_kernel<R, Field, Solver>, kWideBlock, smem));
  if (per_sm < 1) { set_error("the wide kernel does not fit an SM (state dimension %d)", Field::kDim); return DFX_ERR_UNSUPPORTED; }
  long long blocks = (p.n_traj + kWideBlock / 32 - 1) / (kWideBlock / 32);   // one warp per trajectory at a time
  const long long resident = (long long)sms * per_sm;
  if (blocks > resident) blocks = resident;              // persistent: the work queue feeds the resident warps
  unsigned long long *counter = nullptr;
  DFX_CUDA_OK(cudaMallocAsync((void **)&counter, 16, stream));
  if (cudaMemsetAsync(counter, 0, 16, stream) != cudaSuccess ||
      (p.totals && cudaMemsetAsync(p.totals, 0, 4 * sizeof(long long), stream) != cudaSuccess)) {
    cudaFreeAsync(counter, stream);
    set_error("cudaMemsetAsync failed");
    return DFX_ERR_CUDA;
  }
  p.work_counter = counter;
  wide_kernel<R, Field, Solver><<<(unsigned)blocks, kWideBlock, smem, stream>>>(p, fp);
  count_launch();
  const cudaError_t le = cudaGetLastError();
  cudaFreeAsync(counter, stream);
  if (le != cudaSuccess) { set_error("wide_kernel launch failed: %s", cudaGetErrorString(le)); return DFX_ERR_CUDA; }
  return 0;
}

template <class R, class Field, class Solver>
struct WideRegistrar {
  WideRegistrar() { register_builtin(Field::kId, Field::kDim, Solver::kId, DtypeOf<R>::value, 0, &launch_wide<R, Field, Solver>); }
};
#define DFX_REGISTER_WIDE(R, Field, Solver) static ::dfx::WideRegistrar<R, Field, Solver> DFX_CAT(dfx_wide_registrar_, __COUNTER__);

}  // namespace dfx
