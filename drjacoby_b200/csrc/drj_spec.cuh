// drj_spec.cuh -- the SPEC sweep kernel of drjacoby_b200 (sm_100a): one warp per chain, five updates in flight.
//
// What it is for: the reference's everyday shape -- a handful of chains, one rung, a few parameters, a short data
// vector (run_mcmc defaults: 5 chains, 1 rung, R/main.R:213-216; BASELINE configs[0]: ONE chain over n = 100).  A
// Metropolis-Hastings chain is a serial recurrence: update w + 1 starts from the outcome of update w.  One thread per
// chain runs it as one dependent instruction stream (~1100 instructions, 2.9 us per update on a B200: 3.5e5
// updates/s, a third of one CPU core), and dividing ONE update over the lanes of a warp (drj_lanes.cuh) still leaves
// the transcendental functions of that update in series (2.0 us).
//
// Here the warp runs the next FIVE updates of its chain speculatively ("prefetching" Metropolis-Hastings): the
// outcomes of updates w .. w + 4 form a binary tree of 1 + 2 + 4 + 8 + 16 = 31 nodes -- node n at depth k = the
// proposal of update w + k GIVEN the accept / reject outcomes of the k updates before it (the bits of n below its
// leading one) -- and lane n evaluates node n.  What a node needs from its ancestors is arithmetic only:
//   * the random numbers of update w + k depend on nothing (counter-based Philox, the same counters as everywhere);
//   * the bandwidth after the assumed outcomes is the Robbins-Monro recurrence replayed along the path (one fma per
//     ancestor that touched the same parameter) and ONE exp;
//   * phi' = phi + bandwidth * z needs the phi' of the nearest accepted ancestor on the same parameter: a pass down the
//     levels of the tree by shuffles (no transcendental in it);
//   * theta', the Jacobian logs, the log-likelihood and the log-prior of ALL 31 proposals are then evaluated at the
//     same time, one proposal per lane, each with the code the one-thread-per-particle kernel uses
//     (drj_eval_loglike / M::logprior);
//   * the five decisions are then taken in order (four shuffles and the Metropolis-Hastings test per level), following
//     the path the chain really takes; the other 26 evaluations are discarded.
// The warp advances exactly five updates per pass at the latency of ~one; the arithmetic of every update on the true
// path is the arithmetic drj_sweep_body does, in the same order, so the results are BITWISE those of the
// one-thread-per-particle kernels (tests/test_gpu_spec.py) -- the kernel family is a scheduling choice, not a numerical
// one.  Scope: one rung (no coupling inside the tree), one likelihood block, d <= 4, data resident in shared memory.
#pragma once
#include "drj_sampler.cuh"

#define DRJ_SPEC_H 5     // updates in flight per pass (depth of the tree): 2^5 - 1 = 31 nodes, one per lane
#define DRJ_SPEC_MAXD 4  // parameters (the per-lane path state lives in registers)
#define DRJ_SPEC_NT 256  // threads per CTA = 8 chains

template <class M>
struct DrjSpecTraits {
  static constexpr bool ok = (M::NBLOCK == 1) && (M::D <= DRJ_SPEC_MAXD);
};

// element `i` (run-time) of a small register array
template <int D>
__device__ __forceinline__ double drj_pick(const double (&v)[D], int i) {
  double x = v[0];
#pragma unroll
  for (int q = 1; q < D; ++q) x = (i == q) ? v[q] : x;
  return x;
}
template <int D>
__device__ __forceinline__ int drj_picki(const int (&v)[D], int i) {
  int x = v[0];
#pragma unroll
  for (int q = 1; q < D; ++q) x = (i == q) ? v[q] : x;
  return x;
}
__device__ __forceinline__ double drj_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

template <class M>
__device__ __forceinline__ void drj_spec_sweep_body(const DrjKernelArgs& a) {
  constexpr int D = M::D, H = DRJ_SPEC_H;
  static_assert(M::NBLOCK == 1 && D <= DRJ_SPEC_MAXD, "SPEC kernel: one block, d <= 4");
  const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
  const long long chain = (long long)blockIdx.x * a.chains_per_cta + wq;
  const bool valid = (wq < a.chains_per_cta) && (chain < a.chains);
  const long long cc = valid ? chain : 0;  // warps without a chain shadow chain 0 (they take part in the data staging)
  const long long p = cc;                  // one rung: particle = chain
  const long long P = a.P;

  __shared__ double s_tmin[D], s_tmax[D];
  __shared__ int s_trans[D], s_fl[D];
  __shared__ int s_nfree;
  if (tid == 0) {
    int f = 0;
    for (int i = 0; i < D; ++i) {
      s_tmin[i] = a.theta_min[i]; s_tmax[i] = a.theta_max[i]; s_trans[i] = a.trans_type[i];
      if (a.free_slot[i] >= 0) s_fl[f++] = i;  // the free parameters in update order
    }
    s_nfree = f;
  }
  DrjTiles<M> T;
  T.template init<DRJ_MODE_SOLO>(a);  // per-datum models: the data item into shared memory (resident: n <= 2 tiles)
  __syncthreads();
  const int F = s_nfree;

  // ---- the chain's state, replicated in the registers of every lane
  double theta[D], phi[D], logbw[D], alo[D], ahi[D];
  int bwi[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const long long o = (long long)i * P + p;
    theta[i] = a.theta[o]; phi[i] = a.phi[o]; logbw[i] = a.logbw[o];
    alo[i] = a.adj_lo[o]; ahi[i] = a.adj_hi[o]; bwi[i] = a.bw_index[o];
  }
  double ll = a.ll[p], lp = a.lp[p];
  int acc = a.accept[p];
  const double beta = a.beta_raised[0];
  const double target = a.target_acceptance;
  const unsigned gchain_lo = (unsigned)(unsigned long long)(a.chain_offset + cc);
  const unsigned gchain_hi = (unsigned)((unsigned long long)(a.chain_offset + cc) >> 32);
  const unsigned k0 = (unsigned)a.seed, k1 = (unsigned)(a.seed >> 32);

  // this lane's node: n = lane (lane 0 doubles node 1), depth k, outcome of the ancestor at depth j = bit (k-1-j) of n
  const int n = lane ? lane : 1;
  const int k = 31 - __clz(n);
  const int parent = n >> 1;

  const long long V_total = F > 0 ? (long long)a.n_iter * F : 0;  // updates in this launch; update v = iteration * F + free slot
  int it = 0, f0 = 0;  // iteration-in-launch and free slot of the next update
  bool stopped = false;
  for (long long v0 = 0; v0 < V_total && !stopped; v0 += H) {
    // ================= node (lane): the proposal of update v0 + k on the assumed path =================
    const long long vk = v0 + k;
    const int itk = it + (f0 + k) / F, fk = (f0 + k) % F;
    const int ik = s_fl[fk];
    const bool live = vk < V_total;
    const int tk = a.iter0 + (live ? itk : it);
    double u1, u2, u3;
    if (a.rng_mode) {
      const double* qq = a.replay + (cc * a.replay_T + (tk - 1)) * a.replay_U + 3 * (long long)fk;
      u1 = qq[0]; u2 = qq[1]; u3 = qq[2];
    } else {
      DrjU4 x4 = drj_philox4x32_10(gchain_lo, (unsigned)tk, (unsigned)ik, gchain_hi << 8, k0, k1);
      u1 = drj_u32_to_unif(x4.x); u2 = drj_u32_to_unif(x4.y); u3 = drj_u32_to_unif(x4.z);
    }
    const double zq = drj_norm_rand(u1, u2);
    const double lu = log(u3);
    // Robbins-Monro along the path: one step per ancestor that updated this parameter (Particle.h:153-154,165-166)
    double lb = drj_pick<D>(logbw, ik);
    int bi = drj_picki<D>(bwi, ik);
#pragma unroll
    for (int j = 0; j < H - 1; ++j) {
      if (j < k && ((k - j) % F) == 0) {
        const bool aj = (n >> (k - 1 - j)) & 1;
        lb = fma(aj ? (1.0 - target) : -target, rsqrt((double)bi), lb);
        ++bi;
      }
    }
    const double bw = exp(lb);
    // phi of this parameter on the path: down the levels of the tree (a child takes its parent's, with the parent's
    // proposal in place if the path says the parent was accepted)
    double phiC[D];
#pragma unroll
    for (int i = 0; i < D; ++i) phiC[i] = phi[i];
    double phi_p = drj_rnorm_from(drj_pick<D>(phiC, ik), bw, zq);  // (final for the root; deeper levels redo it below)
    const int ipar = s_fl[(fk + F - 1) % F];  // the parameter the parent node proposed
#pragma unroll
    for (int lev = 1; lev < H; ++lev) {
      const double pp = drj_shfl(phi_p, parent);
      double pc[D];
#pragma unroll
      for (int i = 0; i < D; ++i) pc[i] = drj_shfl(phiC[i], parent);
      if (k == lev) {
#pragma unroll
        for (int i = 0; i < D; ++i) phiC[i] = ((n & 1) && i == ipar) ? pp : pc[i];
        phi_p = drj_rnorm_from(drj_pick<D>(phiC, ik), bw, zq);
      }
    }
    // phi_prop_to_theta_prop (Particle.cpp:55-74) and the logs of get_adjustment (Particle.cpp:103-124), all four
    // transform types in one instruction stream (the lanes of a warp propose different parameters)
    const int tt = s_trans[ik];
    const double tmin = s_tmin[ik], tmax = s_tmax[ik];
    double th_p, lo_p, hi_p;
    {
      const double e = exp(phi_p);
      const double t3 = __dadd_rn(__dmul_rn(tmax, e), tmin) / (1.0 + e);  // (no FMA: the reference rounds the product, Particle.cpp:68)
      th_p = (tt == 0) ? phi_p : (tt == 1) ? (tmax - e) : (tt == 2) ? (e + tmin) : t3;
      const double hi_arg = (tt == 1 || tt == 3) ? (tmax - th_p) : 1.0;  // log(1) = +0: the unused log is the reference's 0
      const double lo_arg = (tt == 2 || tt == 3) ? (th_p - tmin) : 1.0;
      hi_p = log(hi_arg);
      lo_p = log(lo_arg);
    }
    // theta and the Jacobian logs of the current state on the path: the same pass down the levels
    double thC[D], loC[D], hiC[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { thC[i] = theta[i]; loC[i] = alo[i]; hiC[i] = ahi[i]; }
#pragma unroll
    for (int lev = 1; lev < H; ++lev) {
      const double pt = drj_shfl(th_p, parent), pl = drj_shfl(lo_p, parent), ph = drj_shfl(hi_p, parent);
      double ct[D], cl[D], ch[D];
#pragma unroll
      for (int i = 0; i < D; ++i) { ct[i] = drj_shfl(thC[i], parent); cl[i] = drj_shfl(loC[i], parent); ch[i] = drj_shfl(hiC[i], parent); }
      if (k == lev) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const bool take = (n & 1) && i == ipar;
          thC[i] = take ? pt : ct[i]; loC[i] = take ? pl : cl[i]; hiC[i] = take ? ph : ch[i];
        }
      }
    }
    double adj;
    {
      const double alo_cur = drj_pick<D>(loC, ik), ahi_cur = drj_pick<D>(hiC, ik);
      adj = (tt == 0) ? 0.0 : (tt == 1) ? (hi_p - ahi_cur) : (tt == 2) ? (lo_p - alo_cur) : (hi_p + lo_p - ahi_cur - alo_cur);
    }
    // likelihood and prior at theta' (Particle.h:107-118), the one-thread-per-particle evaluation
    double thn[D];
#pragma unroll
    for (int i = 0; i < D; ++i) thn[i] = (i == ik) ? th_p : thC[i];
    const double ll_p = drj_eval_loglike<M, DRJ_MODE_SOLO>(T, a, thn, 1, 0);
    const double lp_p = M::logprior(thn, a.misc);

    // ================= the decisions, in order, along the path the chain takes =================
    int c = 1;  // node of the next update
    int err = 0;
#pragma unroll 1
    for (int lev = 0; lev < H; ++lev) {
      if (v0 + lev >= V_total) break;
      const int iu = s_fl[f0];  // (uniform) parameter of this update
      const double n_ll = drj_shfl(ll_p, c), n_lp = drj_shfl(lp_p, c), n_adj = drj_shfl(adj, c), n_lu = drj_shfl(lu, c);
      // NaN / +Inf guard (Particle.h:121-128)
      if (err == 0 && valid) {
        if (isnan(n_ll) || n_ll == DRJ_INF) err = DRJ_ERR_LOGLIKE_BAD;
        else if (isnan(n_lp) || n_lp == DRJ_INF) err = DRJ_ERR_LOGPRIOR_BAD;
        if (err && lane == c) drj_report(a, err, it, chain, 0, iu, p, thn, D);
      }
      // Metropolis-Hastings (Particle.h:131-139): explicitly rounded product and sums, left to right
      double MH;
      if (beta == 0.0) MH = __dadd_rn(n_lp - lp, n_adj);
      else MH = __dadd_rn(__dadd_rn(__dmul_rn(beta, n_ll - ll), n_lp - lp), n_adj);
      const bool accept = (n_lu < MH) && (err == 0);
      // the node's proposal becomes the state; Robbins-Monro on the log scale (Particle.h:153-154,165-166)
      const double n_phi = drj_shfl(phi_p, c), n_th = drj_shfl(th_p, c), n_lo = drj_shfl(lo_p, c), n_hi = drj_shfl(hi_p, c);
#pragma unroll
      for (int i = 0; i < D; ++i) {
        if (i == iu) {
          const double gain = rsqrt((double)bwi[i]);
          logbw[i] = fma(accept ? (1.0 - target) : -target, gain, logbw[i]);
          bwi[i]++;
          if (accept) { phi[i] = n_phi; theta[i] = n_th; alo[i] = n_lo; ahi[i] = n_hi; }
        }
      }
      ll = accept ? n_ll : ll;
      lp = accept ? n_lp : lp;
      acc += accept ? 1 : 0;
      c = 2 * c + (accept ? 1 : 0);
      if (++f0 == F) {
        f0 = 0;
        // ---- end of an iteration: store (main.cpp:165-173), abort on error (Rcpp::stop in the reference)
        if (valid && lane == 0) {
          a.ring_ll[(long long)it * a.chains + chain] = ll;  // (one rung: store_pt and save_hot_draws address the same slots)
          a.ring_lp[(long long)it * a.chains + chain] = lp;
#pragma unroll
          for (int i = 0; i < D; ++i) a.ring_theta[((long long)it * D + i) * a.chains + chain] = theta[i];
        }
        int stop = err;
        if (lane == 0 && *((volatile unsigned long long*)a.err_key) != ~0ULL) stop = 1;
        ++it;
        if (__any_sync(0xffffffffu, stop)) { stopped = true; break; }
      }
    }
  }

  // ---- write state back
  if (valid && lane == 0) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const long long o = (long long)i * P + p;
      a.theta[o] = theta[i]; a.phi[o] = phi[i]; a.logbw[o] = logbw[i];
      a.adj_lo[o] = alo[i]; a.adj_hi[o] = ahi[i]; a.bw_index[o] = bwi[i];
    }
    a.ll_block[p] = ll;
    a.ll[p] = ll; a.lp[p] = lp; a.accept[p] = acc;
  }
}

template <class M>
__global__ void __launch_bounds__(DRJ_SPEC_NT) drj_sweep_kernel_spec(const __grid_constant__ DrjKernelArgs a) {
  drj_spec_sweep_body<M>(a);
}
template <class M, bool ON = DrjSpecTraits<M>::ok>
struct DrjSpecKernel {
  static const void* sweep() { return nullptr; }
};
template <class M>
struct DrjSpecKernel<M, true> {
  static const void* sweep() { return (const void*)drj_sweep_kernel_spec<M>; }
};
