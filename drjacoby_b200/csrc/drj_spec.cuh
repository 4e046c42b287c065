// drj_spec.cuh -- the SPEC sweep kernels of drjacoby_b200 (sm_100a): one CTA per chain, five or seven updates in flight.
//
// What it is for: the reference's everyday shape -- a handful of chains, one rung, a few parameters, a short data
// vector (run_mcmc defaults: 5 chains, 1 rung, R/main.R:213-216; BASELINE configs[0]: ONE chain over n = 100).  A
// Metropolis-Hastings chain is a serial recurrence: update w + 1 starts from the outcome of update w.  One thread per
// chain runs it as one dependent instruction stream (~1100 instructions, 2.9 us per update on a B200: 3.5e5
// updates/s, a third of one CPU core), and dividing ONE update over the lanes of a warp (drj_lanes.cuh) still leaves
// the transcendental functions of that update in series (2.0 us).
//
// Here a CTA runs the next H updates of its chain speculatively ("prefetching" Metropolis-Hastings): the outcomes
// of updates w .. w + H - 1 form a binary tree of 2^H - 1 nodes -- node n at depth k = the proposal of update w + k GIVEN
// the accept / reject outcomes of the k updates before it (the bits of n below its leading one) -- and thread n
// evaluates node n (H = 5: one warp, 31 nodes; H = 7: four warps, 127 nodes, one warp per scheduler of the SM).  Nodes
// talk through small shared-memory tables indexed by node.  What a node needs from its ancestors is arithmetic only:
//   * the random numbers of update w + k depend on nothing (counter-based Philox, the same counters as everywhere);
//   * the bandwidth after the assumed outcomes is the Robbins-Monro recurrence replayed along the path (one fma per
//     ancestor that touched the same parameter) and ONE exp;
//   * phi' = phi + bandwidth * z needs the phi' of the nearest accepted ancestor on the same parameter: a pass down the
//     levels of the tree (no transcendental in it);
//   * theta', the Jacobian logs, the log-likelihood and the log-prior of ALL proposals are then evaluated at the
//     same time, one proposal per thread, each with the code the one-thread-per-particle kernel uses
//     (drj_eval_loglike / M::logprior);
//   * every node then takes its own accept / reject decision, assuming its path (the current log-likelihood on the path is
//     that of its nearest accepted ancestor), ballots collect the outcomes, and the path the chain really takes is
//     H steps of integer arithmetic on that mask; its H proposals are committed, the others discarded.
// The chain advances exactly H updates per pass at the latency of ~one; the arithmetic of every update on the true
// path is the arithmetic drj_sweep_body does, in the same order, so the results are BITWISE those of the
// one-thread-per-particle kernels (tests/test_gpu_spec.py) -- the kernel family is a scheduling choice, not a numerical
// one.  Scope: one rung (no coupling inside the tree), one likelihood block, d <= 4, data resident in shared memory.
#pragma once
#include "drj_sampler.cuh"

#define DRJ_SPEC_MAXD 4  // parameters (the chain's state is replicated in the registers of every thread)
#ifndef DRJ_SPEC_DEEP
#define DRJ_SPEC_DEEP 7  // updates in flight of the deep variant (2^7 threads = four warps per chain, one per scheduler)
#endif

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

template <class M, int H>
__device__ __forceinline__ void drj_spec_sweep_body(const DrjKernelArgs& a) {
  DrjCtaInit<M>::run(a);
  constexpr int D = M::D, NN = 1 << H, NW = NN / 32;  // threads = nodes 1 .. NN-1 (thread 0 doubles node 1), warps
  static_assert(M::NBLOCK == 1 && D <= DRJ_SPEC_MAXD, "SPEC kernels: one block, d <= 4");
  static_assert(H >= 5 && H <= 8, "SPEC kernels: 5 to 8 updates in flight");
  const int tid = threadIdx.x, lane = tid & 31;
  const long long chain = blockIdx.x;  // one chain per CTA, one rung: particle = chain
  const long long p = chain;
  const long long P = a.P;

  __shared__ double s_tmin[D], s_tmax[D];
  __shared__ int s_trans[D], s_fl[D];
  __shared__ int s_nfree;
  // what the nodes publish for each other, by node
  __shared__ double t_phi[NN], t_th[NN], t_lo[NN], t_hi[NN], t_ll[NN], t_lp[NN], t_gain[NN];
  __shared__ int t_err[NN];
  __shared__ unsigned m_acc[NW], m_err[NW];
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

  // ---- the chain's state, replicated in the registers of every thread
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
  const unsigned gchain_lo = (unsigned)(unsigned long long)(a.chain_offset + chain);
  const unsigned gchain_hi = (unsigned)((unsigned long long)(a.chain_offset + chain) >> 32);
  const unsigned k0 = (unsigned)a.seed, k1 = (unsigned)(a.seed >> 32);

  // this thread's node: n = tid (thread 0 doubles node 1), depth k, outcome of the ancestor at depth j = bit (k-1-j) of n
  const int n = tid ? tid : 1;
  const int k = 31 - __clz(n);
  // What the node needs to know about its path never changes (the tree has the same shape in every pass):
  //   rm_mask    bit j set = the ancestor at depth j proposed the node's own parameter (F free parameters in turn:
  //              (k - j) % F == 0): a Robbins-Monro step of that parameter's bandwidth lies on the path;
  //   anc[dl]    node of the nearest ACCEPTED ancestor that proposed the parameter `dl` update slots before the node's
  //              ((k - j) % F == dl), -1 = none: that parameter still has the value of the chain's state;
  //   anc_any    node of the nearest accepted ancestor of any parameter (-1 = none): whose log-likelihood and
  //              log-prior are the current ones on the path.
  unsigned rm_mask = 0u;
  int anc_any = -1;
  int anc[D];
#pragma unroll
  for (int dl = 0; dl < D; ++dl) anc[dl] = -1;
  if (F > 0) {
    for (int j = 0; j < k; ++j) {  // ascending: the nearest ancestor overwrites
      const int dl = (k - j) % F;
      if (dl == 0) rm_mask |= 1u << j;
      if ((n >> (k - 1 - j)) & 1) {
        anc_any = n >> (k - j);
#pragma unroll
        for (int q = 0; q < D; ++q) if (q == dl) anc[q] = n >> (k - j);
      }
    }
  }
  // free slot and iteration offset of update (f0 + k) for every possible first slot f0 < F <= 4, one byte each
  unsigned fk_tab = 0u, it_tab = 0u;
  for (int f = 0; f < F; ++f) {
    fk_tab |= (unsigned)((f + k) % F) << (8 * f);
    it_tab |= (unsigned)((f + k) / F) << (8 * f);
  }

  const long long V_total = F > 0 ? (long long)a.n_iter * F : 0;  // updates in this launch; update v = iteration * F + free slot
  int it = 0, f0 = 0;  // iteration-in-launch and free slot of the next update
  int err = 0;         // first error of the current iteration (the run stops at the iteration's end)
  for (long long v0 = 0; v0 < V_total; v0 += H) {
    // an error reported by another chain stops this one at its next iteration boundary; the flag is read here, a pass
    // early, so that the load's latency is not on the chain's critical path
    unsigned long long seen_key = ~0ULL;
    if (tid == 0) seen_key = *((volatile unsigned long long*)a.err_key);
    // ================= node (thread): the proposal of update v0 + k on the assumed path =================
    const int fk = (int)((fk_tab >> (8 * f0)) & 0xffu), itk = it + (int)((it_tab >> (8 * f0)) & 0xffu);
    const int ik = s_fl[fk];
    const int tk = a.iter0 + ((v0 + k < V_total) ? itk : it);  // (nodes past the end of the launch: any valid row)
    double u1, u2, u3;
    if (a.rng_mode) {
      const double* qq = a.replay + (chain * a.replay_T + (tk - 1)) * a.replay_U + 3 * (long long)fk;
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
      if ((rm_mask >> j) & 1u) {
        const bool aj = (n >> (k - 1 - j)) & 1;
        lb = fma(aj ? (1.0 - target) : -target, rsqrt((double)bi), lb);
        ++bi;
      }
    }
    const double bw = exp(lb);
    const double gain = rsqrt((double)bi);  // of this node's own update
    // phi' = phi + bandwidth * z, phi = the proposal of the nearest accepted ancestor on this parameter (down the levels
    // of the tree: that ancestor's own phi' is final, and published, one level earlier) or the chain's
    const int a0 = anc[0];
    double phi_p = drj_rnorm_from(drj_pick<D>(phi, ik), bw, zq);  // (final when there is no such ancestor)
#pragma unroll  // (rolled, this loop and the commit loop below make the kernel 17 % smaller and 3-8 % slower, same-box A/B)
    for (int lev = 1; lev < H; ++lev) {
      if (k == lev - 1) t_phi[n] = phi_p;
      __syncthreads();
      if (k == lev && a0 >= 0) phi_p = drj_rnorm_from(t_phi[a0], bw, zq);
    }
    // phi_prop_to_theta_prop (Particle.cpp:55-74) and the logs of get_adjustment (Particle.cpp:103-124), all four
    // transform types in one instruction stream (the threads of a warp propose different parameters)
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
    if (k == H - 1) t_phi[n] = phi_p;
    t_th[n] = th_p; t_lo[n] = lo_p; t_hi[n] = hi_p; t_gain[n] = gain;
    __syncthreads();
    // the Jacobian logs of the current state of this parameter on the path, and the other parameters' values there
    double adj;
    {
      const double alo_cur = a0 < 0 ? drj_pick<D>(alo, ik) : t_lo[a0], ahi_cur = a0 < 0 ? drj_pick<D>(ahi, ik) : t_hi[a0];
      adj = (tt == 0) ? 0.0 : (tt == 1) ? (hi_p - ahi_cur) : (tt == 2) ? (lo_p - alo_cur) : (hi_p + lo_p - ahi_cur - alo_cur);
    }
    double thn[D], lon[D];  // (lon: the Jacobian logarithms log(theta_i - min_i) that go with thn, for models with HINTS)
#pragma unroll
    for (int i = 0; i < D; ++i) { thn[i] = (i == ik) ? th_p : theta[i]; lon[i] = (i == ik) ? lo_p : alo[i]; }
#pragma unroll
    for (int dl = 1; dl < D; ++dl) {
      const int src = anc[dl];
      if (dl < F && src >= 0) {
        const double pt = t_th[src], pl = t_lo[src];
        const int fo = fk - dl + (fk < dl ? F : 0);  // the free slot `dl` updates before this node's
        const int io = s_fl[fo];
#pragma unroll
        for (int i = 0; i < D; ++i) { thn[i] = (i == io) ? pt : thn[i]; lon[i] = (i == io) ? pl : lon[i]; }
      }
    }
    // likelihood and prior at theta' (Particle.h:107-118), the one-thread-per-particle evaluation
    const double ll_p = drj_eval_loglike<M, DRJ_MODE_SOLO>(T, a, thn, 1, 0, lon);
    const double lp_p = DrjHints<M>::logprior(thn, a.misc, lon, a.lo_is_log);

    // ================= the decisions =================
    // Every node takes ITS OWN decision at once, assuming its path: the current log-likelihood / log-prior on the path
    // are those of its nearest accepted ancestor (a fixed node) or the chain's.  Ballots collect the outcomes; the
    // path the chain really takes is then H steps of integer arithmetic on that mask (no floating point in the serial
    // part), and on the true path every assumption held, so the decisions are the serial ones.
    int n_err = 0;  // NaN / +Inf guard (Particle.h:121-128)
    if (isnan(ll_p) || ll_p == DRJ_INF) n_err = DRJ_ERR_LOGLIKE_BAD;
    else if (isnan(lp_p) || lp_p == DRJ_INF) n_err = DRJ_ERR_LOGPRIOR_BAD;
    t_ll[n] = ll_p; t_lp[n] = lp_p; t_err[n] = n_err;
    __syncthreads();
    bool n_acc;
    {
      const double ll_cur = anc_any < 0 ? ll : t_ll[anc_any], lp_cur = anc_any < 0 ? lp : t_lp[anc_any];
      // Metropolis-Hastings (Particle.h:131-139): explicitly rounded product and sums, left to right
      double MH;
      if (beta == 0.0) MH = __dadd_rn(lp_p - lp_cur, adj);
      else MH = __dadd_rn(__dadd_rn(__dmul_rn(beta, ll_p - ll_cur), lp_p - lp_cur), adj);
      n_acc = (lu < MH) && (n_err == 0);
    }
    {
      const unsigned ba = __ballot_sync(0xffffffffu, n_acc), be = __ballot_sync(0xffffffffu, n_err != 0);
      if (lane == 0) { m_acc[tid >> 5] = ba; m_err[tid >> 5] = be; }
    }
    __syncthreads();
    unsigned wa[NW], we[NW];
#pragma unroll
    for (int q = 0; q < NW; ++q) { wa[q] = m_acc[q]; we[q] = m_err[q]; }
    const int nlev = (V_total - v0 < H) ? (int)(V_total - v0) : H;
    int path[H];
    unsigned bits = 0u;
    int err_lev = -1;
    {
      int c = 1;
#pragma unroll
      for (int lev = 0; lev < H; ++lev) {
        path[lev] = c;
        unsigned ca = wa[0], ce = we[0];
#pragma unroll
        for (int q = 1; q < NW; ++q) { ca = ((c >> 5) == q) ? wa[q] : ca; ce = ((c >> 5) == q) ? we[q] : ce; }
        if (err == 0 && err_lev < 0 && ((ce >> (c & 31)) & 1u)) err_lev = lev;  // the first bad proposal on the true path
        // (after an error every update of the iteration is rejected, and the run stops at the iteration's end)
        const unsigned ok = ((ca >> (c & 31)) & 1u) & ((err == 0 && err_lev < 0) ? 1u : 0u);
        bits |= ok << lev;
        c = 2 * c + (int)ok;
      }
    }
    // ================= commit, level by level =================
    bool stopped = false;
#pragma unroll
    for (int lev = 0; lev < H; ++lev) {
      if (lev < nlev && !stopped) {
        const int c = path[lev];
        const bool accept = (bits >> lev) & 1u;
        const int iu = s_fl[f0];  // (uniform) parameter of this update
        if (lev == err_lev) {
          err = t_err[c];
          if (tid == c) drj_report(a, n_err, it, chain, 0, iu, p, thn, D);
        }
        // the node's proposal becomes the state; Robbins-Monro on the log scale (Particle.h:153-154,165-166)
        const double n_gain = t_gain[c];  // rsqrt(bandwidth index) as the node computed it on this very path
#pragma unroll
        for (int i = 0; i < D; ++i) {
          if (i == iu) {
            logbw[i] = fma(accept ? (1.0 - target) : -target, n_gain, logbw[i]);
            bwi[i]++;
            if (accept) { phi[i] = t_phi[c]; theta[i] = t_th[c]; alo[i] = t_lo[c]; ahi[i] = t_hi[c]; }
          }
        }
        if (accept) { ll = t_ll[c]; lp = t_lp[c]; }
        acc += accept ? 1 : 0;
        if (++f0 == F) {
          f0 = 0;
          // ---- end of an iteration: store (main.cpp:165-173), abort on error (Rcpp::stop in the reference)
          if (tid == 0) {
            a.ring_ll[(long long)it * a.chains + chain] = ll;  // (one rung: store_pt and save_hot_draws address the same slots)
            a.ring_lp[(long long)it * a.chains + chain] = lp;
#pragma unroll
            for (int i = 0; i < D; ++i) a.ring_theta[((long long)it * D + i) * a.chains + chain] = theta[i];
          }
          ++it;
          if (err) stopped = true;
        }
      }
    }
    // the tables are rewritten by the next pass; an error seen elsewhere (thread 0) or here stops the chain
    if (__syncthreads_or((stopped || (seen_key != ~0ULL && f0 == 0)) ? 1 : 0)) break;
  }

  // ---- write state back
  if (tid == 0) {
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

template <class M, int H>
__global__ void __launch_bounds__(1 << H) drj_sweep_kernel_spec(const __grid_constant__ DrjKernelArgs a) {
  drj_spec_sweep_body<M, H>(a);
}
template <class M, bool ON = DrjSpecTraits<M>::ok>
struct DrjSpecKernel {
  static const void* sweep5() { return nullptr; }
  static const void* sweep7() { return nullptr; }
};
template <class M>
struct DrjSpecKernel<M, true> {
  static const void* sweep5() { return (const void*)drj_sweep_kernel_spec<M, 5>; }
  static const void* sweep7() { return (const void*)drj_sweep_kernel_spec<M, DRJ_SPEC_DEEP>; }
};
