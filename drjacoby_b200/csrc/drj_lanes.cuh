// drj_lanes.cuh -- the LANES sweep kernel of drjacoby_b200 (sm_100a): a group of L lanes per particle.
//
// Same update as drj_sweep_body (Particle::update src/Particle.h:82-171, coupling src/main.cpp:282-344, storage
// src/main.cpp:165-173), for the two shapes where one THREAD per particle is the wrong grain:
//
//   * many parameters (d > 8; K4: 50 parameters in 51 overlapping blocks).  With one thread per particle the state
//     (5 d + 2 B doubles, 3 KB at d = 50) lives in local memory, 512 resident threads x 3 KB thrash L1 and L2
//     (round-2 ncu of the d = 50 model: 60 GB of DRAM reads + 60 GB of writes per 134 ms launch -- LOCAL memory
//     spilling to HBM --, long_scoreboard 5.7 stalls per issue, FP64 pipe 18 %), and every update walks the 50-term
//     block that ties the parameters together serially.
//   * few particles over a short data item (the reference's everyday shape: run_mcmc defaults to 5 chains,
//     R/main.R:213-216; BASELINE configs[0] is ONE chain over n = 100).  One thread per particle is a single
//     dependent instruction chain of ~1100 instructions per update (2.9 us: Philox -> qnorm -> exp -> log -> the
//     100-term sum -> ...), slower than one CPU core.
//
// Here a particle is owned by a group of L lanes of one warp (L = 8 for many-parameter models, L = 32 for per-datum
// models).  Its state lives ONCE in shared memory (not once per thread): L-fold less on-chip state per particle, no
// local memory.  Per update:
//   * the part that does not depend on the chain's state -- Philox, the uniforms, qnorm (AS241) of the proposal,
//     log u of the acceptance test, and (when d >= L) exp(log bandwidth), rsqrt(bandwidth index) -- is computed for L
//     consecutive updates AT ONCE, one update per lane, and handed to the serial part by shuffles: the longest
//     stretch of the per-update dependent chain is paid once per L updates (for d < L a batch spans several
//     iterations: counter-based random numbers do not depend on the order they are drawn in);
//   * the likelihood of a block is divided over the lanes (a per-datum model: lane l takes the data l, l + L, ...;
//     a generic model opts in with `LANES` / `loglike_lanes`, its own split) and summed by an xor-shuffle tree;
//     the index-order sum over the block likelihoods (misc.h:20-27) is divided the same way.  Association depends
//     only on L (a property of the model) and n: results are bitwise independent of the launch geometry, the
//     launch split and the sharding over GPUs, and agree with the one-thread-per-particle kernels inside the
//     1e-10 gate;
//   * everything else (transform, Jacobian, Metropolis-Hastings, Robbins-Monro) is computed redundantly by the L
//     lanes (deterministic arithmetic: the replicas agree bit for bit), lane 0 writes;
//   * Metropolis coupling swaps two replicas by swapping the INDICES of their shared-memory state blocks (rung-owned
//     state -- bandwidths, counters -- is a separate block that stays), not by copying 4 d + B + 2 doubles.
// Random numbers are the same counters as in drj_sweep_body, so a run is the same chain whichever kernel family
// executes it (up to the association of the likelihood sums).
#pragma once
#include "drj_sampler.cuh"

#define DRJ_MODE_LANES 3

// A generic model opts in to the lane-split form with
//   static constexpr int LANES = 8;   // lanes per particle: 2, 4, 8, 16 or 32
//   __device__ static double loglike_lanes(const double* theta, const DrjList& data, const DrjList& misc, int block, int lane);
// = the terms of loglike(theta, data, misc, block) that lane `lane` of LANES adds up (their sum over the lanes is the
// block's log-likelihood).  Per-datum models need nothing: their split is lane l -> data l, l + 32, ...
template <class M, class = void>
struct DrjLaneForm {
  static constexpr bool generic = false;
  static constexpr int L = 32;
};
template <class M>
struct DrjLaneForm<M, decltype((void)M::LANES)> {
  static constexpr bool generic = true;
  static constexpr int L = M::LANES;
};
template <class M>
struct DrjLanesTraits {
  static constexpr bool ok = M::DATUM_FORM || DrjLaneForm<M>::generic;
  static constexpr int L = DrjLaneForm<M>::L;
  static constexpr int D = M::D, B = M::NBLOCK;
  static constexpr bool BATCH = D >= L;  // a batch of L consecutive updates touches every parameter at most once
  // shared-memory blocks of a particle, in doubles (odd strides: the groups of a warp fall in distinct banks)
  static constexpr int RBS = (4 * D + 2 * B + 2) | 1;      // replica-owned: theta, phi, alo, ahi [D]; llb [B]; the undo list of an update [B]
  static constexpr int RGS = (BATCH ? D : 3 * D) | 1;   // rung-owned: logbw [D] (+ bw, gain [D] when not batched); + bwi [D] ints
  static constexpr int bytes_per_particle = (RBS + RGS) * 8 + D * 4;
};

template <int L>
__device__ __forceinline__ double drj_group_sum(double v) {
#pragma unroll
  for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // (a + b == b + a: every lane of the group holds the same bits)
}
template <int L>
__device__ __forceinline__ double drj_group_get(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src, L);
}

// log-likelihood of `block` at `theta` (shared memory), evaluated by the L lanes of the group; every lane returns it
template <class M>
__device__ __forceinline__ double drj_lanes_block(const double* theta, const DrjList& data, const DrjList& misc, int block, int lane) {
  constexpr int L = DrjLanesTraits<M>::L;
  if constexpr (M::DATUM_FORM) {
    typename M::Consts k;
    const bool fast = M::prepare(theta, misc, block, k);
    const int item = M::datum_item(block);
    const long long n = data.len(item);
    const double* __restrict__ x = data.item(item);
    // lane l: data l, l + L, ... in order, dealt to four partial sums (a short item is a single pass)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long j = lane;
    for (; j + 3 * L < n; j += 4 * L) {
      a0 = M::datum(k, x[j], a0);
      a1 = M::datum(k, x[j + L], a1);
      a2 = M::datum(k, x[j + 2 * L], a2);
      a3 = M::datum(k, x[j + 3 * L], a3);
    }
    if (j < n) a0 = M::datum(k, x[j], a0);
    if (j + L < n) a1 = M::datum(k, x[j + L], a1);
    if (j + 2 * L < n) a2 = M::datum(k, x[j + 2 * L], a2);
    const double acc = drj_group_sum<L>((a0 + a1) + (a2 + a3));
    if (fast) return M::finish(k, acc, n);
    return M::loglike(theta, data, misc, block);  // degenerate theta: exact serial evaluation
  } else if constexpr (DrjLaneForm<M>::generic) {
    return drj_group_sum<L>(M::loglike_lanes(theta, data, misc, block, lane));
  } else {
    return M::loglike(theta, data, misc, block);
  }
}

// propose_phi (Particle.cpp:48-50), phi_prop_to_theta_prop and get_adjustment (Particle.cpp:55-74,103-124) of one
// parameter, from the standard normal draw `zq`; alo / ahi = log(theta - min), log(max - theta) of the current state
__device__ __forceinline__ void drj_lanes_propose(double phi_cur, double bw, double zq, int trans, double tmin, double tmax,
                                                  double alo_cur, double ahi_cur, double& phi_p, double& th_p, double& lo_p,
                                                  double& hi_p, double& adj) {
  phi_p = drj_rnorm_from(phi_cur, bw, zq);
  lo_p = 0.0; hi_p = 0.0; adj = 0.0;
  switch (trans) {
    case 0: th_p = phi_p; break;
    case 1: th_p = tmax - exp(phi_p); hi_p = log(tmax - th_p); adj = hi_p - ahi_cur; break;
    case 2: th_p = exp(phi_p) + tmin; lo_p = log(th_p - tmin); adj = lo_p - alo_cur; break;
    default: {
      const double e = exp(phi_p);
      th_p = __dadd_rn(__dmul_rn(tmax, e), tmin) / (1.0 + e);  // (no FMA: the reference rounds the product, Particle.cpp:68)
      hi_p = log(tmax - th_p); lo_p = log(th_p - tmin);
      adj = hi_p + lo_p - ahi_cur - alo_cur;
    } break;
  }
}

// per-CTA tables of the coupling pass, one entry per particle slot
template <class M>
struct DrjLanesShared {
  static constexpr int NPMAX = DRJ_NT / DrjLanesTraits<M>::L;
  double ll[NPMAX], lp[NPMAX], q1[NPMAX], q2[NPMAX], logu[NPMAX];
  int slot[NPMAX];
  unsigned swap[NPMAX];  // one mask per chain (rungs <= NPMAX <= 64 ... at most 32 pairs are used: rungs <= 32)
};

template <class M>
__device__ __forceinline__ void drj_lanes_sweep_body(const DrjKernelArgs& a) {
  DrjCtaInit<M>::run(a);
  typedef DrjLanesTraits<M> TR;
  constexpr int L = TR::L, D = TR::D, B = TR::B, RBS = TR::RBS, RGS = TR::RGS;
  constexpr bool BATCH = TR::BATCH;
  static_assert(L == 2 || L == 4 || L == 8 || L == 16 || L == 32, "LANES must be 2, 4, 8, 16 or 32");
  const int R = a.rungs;
  const int tid = threadIdx.x, lane = tid & (L - 1), q = tid / L;
  const int NP = blockDim.x / L;
  const int lc = q / R, r = q - lc * R;
  const long long chain = (long long)blockIdx.x * a.chains_per_cta + lc;
  const bool valid = (lc < a.chains_per_cta) && (chain < a.chains);
  const bool active = valid && lane == 0;
  const long long cc = valid ? chain : 0;  // groups without a particle shadow chain 0 (uniform control flow)
  const long long p = cc * R + r;
  const long long P = a.P;

  __shared__ double s_tmin[D], s_tmax[D];
  __shared__ int s_trans[D], s_free[D], s_bptr[D + 1];
  __shared__ double s_beta[DRJ_NT / L];
  __shared__ DrjLanesShared<M> sh;
  for (int i = tid; i < R; i += blockDim.x) s_beta[i] = a.beta_raised[i];
  for (int i = tid; i < D; i += blockDim.x) {
    s_tmin[i] = a.theta_min[i]; s_tmax[i] = a.theta_max[i];
    s_trans[i] = a.trans_type[i]; s_free[i] = a.free_slot[i];
  }
  for (int i = tid; i <= D; i += blockDim.x) s_bptr[i] = a.block_ptr[i];

  // ---- this particle's state blocks in shared memory
  double* const RB0 = drj_dyn_smem;                    // [NP][RBS]
  double* const RG0 = RB0 + (size_t)NP * RBS;          // [NP][RGS]
  int* const BWI0 = (int*)(RG0 + (size_t)NP * RGS);    // [NP][D]
  int* const s_bidx = BWI0 + (size_t)NP * D;           // the parameters' block lists (1-based), a.block_ptr[D] entries
  for (int i = tid, n = a.block_ptr[D]; i < n; i += blockDim.x) s_bidx[i] = a.block_idx[i];
  const DrjList& data = a.data;
  int myslot = q;                                      // replica block this rung holds (changes when replicas swap)
  double* const logbw = RG0 + (size_t)q * RGS;
  double* const bwc = logbw + D;                       // (only when !BATCH)
  double* const gainc = logbw + 2 * D;
  int* const bwi = BWI0 + (size_t)q * D;
  {
    double* rb = RB0 + (size_t)myslot * RBS;
    for (int i = lane; i < D; i += L) {
      const long long o = (long long)i * P + p;
      rb[i] = a.theta[o]; rb[D + i] = a.phi[o]; rb[2 * D + i] = a.adj_lo[o]; rb[3 * D + i] = a.adj_hi[o];
      const double lb = a.logbw[o];
      const int bi = a.bw_index[o];
      logbw[i] = lb; bwi[i] = bi;
      if (!BATCH) { bwc[i] = exp(lb); gainc[i] = rsqrt((double)bi); }
    }
    for (int b = lane; b < B; b += L) {
      rb[4 * D + b] = a.ll_block[(long long)b * P + p];
    }
  }
  double ll = a.ll[p], lp = a.lp[p];
  int acc = a.accept[p];
  int mc_count = 0;
  const double beta = a.beta_raised[r];
  const double target = a.target_acceptance;
  const unsigned gchain_lo = (unsigned)(unsigned long long)(a.chain_offset + cc);
  const unsigned gchain_hi = (unsigned)((unsigned long long)(a.chain_offset + cc) >> 32);
  const unsigned k0 = (unsigned)a.seed, k1 = (unsigned)(a.seed >> 32);
  const bool do_couple = a.coupling_on && R > 1;
  __syncthreads();

  // the batch: this lane's share of the next L updates.
  //   BATCH (d >= L): the L parameters i .. i + L - 1 of the current iteration.  Each is touched once, so its own
  //   state (phi, bandwidth, Jacobian logs) is current when the batch is formed and the WHOLE proposal -- the draw,
  //   phi', the transform with its exp / division / logs, the Jacobian adjustment, log u -- is computed lane-parallel;
  //   the serial part only broadcasts theta', the adjustment and log u, and the owner lane writes the state.
  //   Otherwise (d < L): the updates w .. w + L - 1 of the launch (w = iteration * d + parameter), over several
  //   iterations; only the part that does not depend on the chain's state (the draw, log u) can be formed ahead.
  double b_zq = 0.0, b_lu = 0.0;                                                  // both modes
  double b_gain = 0.0, b_phi = 0.0, b_th = 0.0, b_lo = 0.0, b_hi = 0.0, b_adj = 0.0, b_thcur = 0.0;  // BATCH
  int it = 0, i = 0, err = 0;
  bool stopped = false;
  long long w = 0;
  while (it < a.n_iter && !stopped) {
    const int j = BATCH ? (i & (L - 1)) : (int)(w & (L - 1));
    if (j == 0) {
      int itl, il;
      bool mine;
      if (BATCH) { itl = it; il = i + lane; mine = il < D; }
      else {
        const long long wl = w + lane;
        itl = (int)(wl / D); il = (int)(wl - (long long)itl * D);
        mine = itl < a.n_iter;
      }
      const int fsl = mine ? s_free[il] : -1;
      if (fsl >= 0) {
        const int tl = a.iter0 + itl;
        double u1, u2, u3;
        if (a.rng_mode) {
          const double* qq = a.replay + (cc * a.replay_T + (tl - 1)) * a.replay_U + 3 * ((long long)r * a.d_free + fsl);
          u1 = qq[0]; u2 = qq[1]; u3 = qq[2];
        } else {
          DrjU4 x4 = drj_philox4x32_10(gchain_lo, (unsigned)tl, ((unsigned)r << 16) | (unsigned)il, gchain_hi << 8, k0, k1);
          u1 = drj_u32_to_unif(x4.x); u2 = drj_u32_to_unif(x4.y); u3 = drj_u32_to_unif(x4.z);
        }
        b_zq = drj_norm_rand(u1, u2);
        b_lu = log(u3);
        if (BATCH) {
          const double* const rb = RB0 + (size_t)myslot * RBS;
          b_gain = rsqrt((double)bwi[il]);
          b_thcur = rb[il];
          drj_lanes_propose(rb[D + il], exp(logbw[il]), b_zq, s_trans[il], s_tmin[il], s_tmax[il], rb[2 * D + il], rb[3 * D + il],
                            b_phi, b_th, b_lo, b_hi, b_adj);
        }
      }
    }
    const int fs = s_free[i];
    if (fs >= 0) {  // (skip_param: no RNG consumed, Particle.h:91-93)
      double* const rb = RB0 + (size_t)myslot * RBS;
      double* const theta = rb;
      double* const llb = rb + 4 * D;
      double* const undo = llb + B;
      const double lu = drj_group_get<L>(b_lu, j);
      double th_p, adj;
      double phi_p = 0.0, lo_p = 0.0, hi_p = 0.0, gain = 0.0, th_cur = 0.0;  // (BATCH: the owner lane has them in b_*)
      if (BATCH) {
        th_p = drj_group_get<L>(b_th, j);
        adj = drj_group_get<L>(b_adj, j);
      } else {
        // propose_phi (Particle.cpp:48-50), phi_prop_to_theta_prop + get_adjustment (Particle.cpp:55-74,103-124)
        const double zq = drj_group_get<L>(b_zq, j);
        gain = gainc[i];
        th_cur = theta[i];
        drj_lanes_propose(rb[D + i], bwc[i], zq, s_trans[i], s_tmin[i], s_tmax[i], rb[2 * D + i], rb[3 * D + i], phi_p, th_p, lo_p, hi_p, adj);
      }
      const bool owner = BATCH ? (lane == j) : (lane == 0);  // the lane that writes this update's state
      __syncwarp();
      if (owner) theta[i] = th_p;  // theta is theta_prop from here to the accept / reject
      __syncwarp();
      // likelihood of the blocks this parameter belongs to (Particle.h:107-112).  The new values go straight into the
      // block table; the owner keeps the old ones in an undo list and puts them back if the proposal is rejected.
      const int jb0 = s_bptr[i], jb1 = s_bptr[i + 1];
      double ll_p, old0 = 0.0, old1 = 0.0;
      int ob0 = 0, ob1 = 0;
      if constexpr (B == 1) {
        ll_p = llb[0];
#pragma unroll 1
        for (int jb = jb0; jb < jb1; ++jb) ll_p = drj_lanes_block<M>(theta, data, a.misc, s_bidx[jb], lane);
      } else {
#pragma unroll 1
        for (int jb = jb0; jb < jb1; ++jb) {
          const int b = s_bidx[jb] - 1;
          const double v = drj_lanes_block<M>(theta, data, a.misc, b + 1, lane);
          if (owner) {  // the first two old values stay in registers, the rest goes to the undo list
            const double o = llb[b];
            if (jb == jb0) { old0 = o; ob0 = b; }
            else if (jb == jb0 + 1) { old1 = o; ob1 = b; }
            else undo[jb - jb0] = o;
            llb[b] = v;
          }
        }
        __syncwarp();
        // index-order sum over the blocks (misc.h:20-27), lane l the blocks l, l + L, ...
        constexpr int NB = (B + L - 1) / L;
        double part = 0.0;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const int b = lane + k * L;
          part += (b < B) ? llb[b] : 0.0;
        }
        ll_p = drj_group_sum<L>(part);
      }
      const double lp_p = M::logprior(theta, a.misc);
      // NaN / +Inf guard (Particle.h:121-128)
      if (err == 0 && valid) {
        if (isnan(ll_p) || ll_p == DRJ_INF) err = DRJ_ERR_LOGLIKE_BAD;
        else if (isnan(lp_p) || lp_p == DRJ_INF) err = DRJ_ERR_LOGPRIOR_BAD;
        if (err && active) drj_report(a, err, it, chain, r, i, p, theta, D);
      }
      // Metropolis-Hastings (Particle.h:131-139): explicitly rounded product and sums, left to right
      double MH;
      if (beta == 0.0) MH = __dadd_rn(lp_p - lp, adj);
      else MH = __dadd_rn(__dadd_rn(__dmul_rn(beta, ll_p - ll), lp_p - lp), adj);
      const bool accept = (lu < MH) && (err == 0);
      __syncwarp();  // every lane is done reading theta' (logprior, error report)
      if (owner) {
        if (BATCH) { phi_p = b_phi; lo_p = b_lo; hi_p = b_hi; gain = b_gain; th_cur = b_thcur; }
        if (accept) {
          rb[D + i] = phi_p; rb[2 * D + i] = lo_p; rb[3 * D + i] = hi_p;
          if constexpr (B == 1) llb[0] = ll_p;
        } else {
          theta[i] = th_cur;
          if constexpr (B > 1) {
            if (jb1 > jb0) llb[ob0] = old0;
            if (jb1 > jb0 + 1) llb[ob1] = old1;
#pragma unroll 1
            for (int jb = jb0 + 2; jb < jb1; ++jb) llb[s_bidx[jb] - 1] = undo[jb - jb0];
          }
        }
        // Robbins-Monro on the log scale: + (1 - a*) / sqrt(index) on acceptance (Particle.h:153), - a* / sqrt(index) otherwise (:165)
        const double lb_new = fma(accept ? (1.0 - target) : -target, gain, logbw[i]);
        logbw[i] = lb_new;
        const int bi = bwi[i] + 1;
        bwi[i] = bi;
        if (!BATCH) { bwc[i] = exp(lb_new); gainc[i] = rsqrt((double)bi); }
      }
      ll = accept ? ll_p : ll;
      lp = accept ? lp_p : lp;
      acc += accept ? 1 : 0;
      __syncwarp();
    }
    ++w;

    if (++i == D) {
      i = 0;
      // ================= store (before coupling, main.cpp:165-173) =================
      if (valid) {
        const double* const theta = RB0 + (size_t)myslot * RBS;
        if (lane == 0) {
          if (a.store_pt) {
            a.ring_ll[(long long)it * P + p] = ll;
            a.ring_lp[(long long)it * P + p] = lp;
          } else if (r == R - 1) {
            a.ring_ll[(long long)it * a.chains + chain] = ll;
            a.ring_lp[(long long)it * a.chains + chain] = lp;
          }
        }
        if (a.save_hot_draws) {
          for (int ii = lane; ii < D; ii += L) a.ring_theta[((long long)it * D + ii) * P + p] = theta[ii];
        } else if (r == R - 1) {
          for (int ii = lane; ii < D; ii += L) a.ring_theta[((long long)it * D + ii) * a.chains + chain] = theta[ii];
        }
      }
      // ================= Metropolis coupling (main.cpp:282-344) =================
      if (do_couple) {
        const int t = a.iter0 + it;
        if (lane == 0) {
          sh.ll[q] = ll; sh.lp[q] = lp; sh.slot[q] = myslot;
          if (r > 0) { sh.q1[q] = __dmul_rn(ll, s_beta[r - 1]); sh.q2[q] = __dmul_rn(ll, beta); }
        }
        if (r < R - 1) {
          double u;
          if (a.rng_mode) u = a.replay[(cc * a.replay_T + (t - 1)) * a.replay_U + 3LL * a.d_free * R + r];
          else {
            DrjU4 x4 = drj_philox4x32_10(gchain_lo, (unsigned)t, (unsigned)r, 1u | (gchain_hi << 8), k0, k1);
            u = drj_u32_to_unif(x4.x);
          }
          const double lgu = log(u);
          if (lane == 0) sh.logu[q] = lgu;
        }
        __syncthreads();
        // sequential part hot -> cold, one walker per chain: decides every pair into a bit mask (see drj_sweep_body)
        const bool whole = lc < a.chains_per_cta;  // (left-over slots of the CTA hold no whole chain)
        if (lane == 0 && r == 0 && whole) {
          const int wb = q;
          unsigned m = 0u;
          double cur_ll = sh.ll[wb];
          double b1 = s_beta[0];
          for (int k = 0; k < R - 1; ++k) {
            const double ll2 = sh.ll[wb + k + 1], q1 = sh.q1[wb + k + 1], q2 = sh.q2[wb + k + 1];
            const double b2 = s_beta[k + 1], lu2 = sh.logu[wb + k];
            // main.cpp:300-308: (loglike2 * beta1 + loglike1 * beta2) - (loglike1 * beta1 + loglike2 * beta2), all rounded
            const double t1 = __dmul_rn(cur_ll, b2), t2 = __dmul_rn(cur_ll, b1);
            const double acceptance = (b1 == 0.0) ? __dsub_rn(t1, q2) : __dsub_rn(__dadd_rn(q1, t1), __dadd_rn(t2, q2));
            const bool swap = lu2 < acceptance;
            m |= (swap ? 1u : 0u) << k;
            cur_ll = swap ? cur_ll : ll2;
            b1 = b2;
          }
          sh.swap[lc] = m;
        }
        __syncthreads();
        if (whole) {
          // rung r receives the replica of rung r + 1 if pair r swapped, else the replica carried up from below the
          // run of swapped pairs that ends at pair r - 1
          const unsigned m = sh.swap[lc];
          int src;
          if (r < R - 1 && ((m >> r) & 1u)) {
            src = r + 1;
            ++mc_count;
          } else {
            const unsigned zeros = ~m & ((1u << r) - 1u);
            src = zeros ? 32 - __clz(zeros) : 0;
          }
          const int from = lc * R + src;
          myslot = sh.slot[from];  // replica-owned state moves by index; rung-owned state (bandwidths, counters) stays
          ll = sh.ll[from];
          lp = sh.lp[from];
        }
        __syncthreads();  // (the tables are rewritten at the next iteration's coupling pass)
      }
      // ================= abort on error (Rcpp::stop in the reference) =================
      int stop = err;
      if (tid == 0 && *((volatile unsigned long long*)a.err_key) != ~0ULL) stop = 1;
      if (__syncthreads_or(stop)) stopped = true;
      ++it;
    }
  }
  __syncthreads();

  // ---- write state back
  if (valid) {
    const double* rb = RB0 + (size_t)myslot * RBS;
    for (int ii = lane; ii < D; ii += L) {
      const long long o = (long long)ii * P + p;
      a.theta[o] = rb[ii]; a.phi[o] = rb[D + ii]; a.adj_lo[o] = rb[2 * D + ii]; a.adj_hi[o] = rb[3 * D + ii];
      a.logbw[o] = logbw[ii]; a.bw_index[o] = bwi[ii];
    }
    for (int b = lane; b < B; b += L) a.ll_block[(long long)b * P + p] = rb[4 * D + b];
    if (lane == 0) {
      a.ll[p] = ll; a.lp[p] = lp; a.accept[p] = acc;
      if (do_couple && r < R - 1) a.mc_accept[chain * (R - 1) + r] += mc_count;
    }
  }
}

template <class M>
__global__ void __launch_bounds__(DRJ_NT, 2) drj_sweep_kernel_lanes(const __grid_constant__ DrjKernelArgs a) {
  drj_lanes_sweep_body<M>(a);
}
template <class M, bool ON = DrjLanesTraits<M>::ok>
struct DrjLanesKernel {
  static const void* sweep() { return nullptr; }
  static constexpr int lanes = 0, bytes_per_particle = 0;
};
template <class M>
struct DrjLanesKernel<M, true> {
  static const void* sweep() { return (const void*)drj_sweep_kernel_lanes<M>; }
  static constexpr int lanes = DrjLanesTraits<M>::L, bytes_per_particle = DrjLanesTraits<M>::bytes_per_particle;
};
