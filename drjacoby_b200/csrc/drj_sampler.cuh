// drj_sampler.cuh -- the MCMC sweep kernels of drjacoby_b200 (sm_100a), templated on a model.
//
// What it replaces (reference, one call per chain, rungs serial):
//   Particle::init / theta_to_phi / init_like    src/Particle.cpp:8-44,78-99, src/Particle.h:59-79
//   Particle::update                             src/Particle.h:82-171
//     propose_phi / phi_prop_to_theta_prop / get_adjustment   src/Particle.cpp:48-74,103-124
//   coupling                                     src/main.cpp:282-344
//   run_mcmc<T1,T2> iteration loop and storage   src/main.cpp:84-278
//
// B200 design (not a port): ONE thread owns ONE (chain, rung) particle and keeps its state in
// registers across all iterations of a launch; all rungs of a chain sit in one CTA so the
// Metropolis-coupling pass is a shared-memory pass; every chain x rung advances in lock-step, so a
// per-datum likelihood is evaluated CTA-wide: data tiles are staged ONCE per CTA into shared
// memory by the TMA bulk-copy engine (cp.async.bulk + mbarrier, double buffered) and every thread
// walks the tile with broadcast LDS, accumulating its own sum (8 interleaved partial sums, assigned by
// global datum index).  Samples go to
// [iteration][particle] rings (coalesced) and are transposed on the device into the reference's
// per-chain layout.
//
// Compiled by nvcc (built-in models) and by NVRTC (user models): no standard headers here.
#pragma once
#include "drj_device.cuh"

#ifndef DRJ_NT
#define DRJ_NT 256      // max threads per CTA
#endif
#ifndef DRJ_TILE
#define DRJ_TILE 2048   // doubles per data tile (16 KB); two tiles are resident per CTA
#endif
#ifndef DRJ_MINB
#define DRJ_MINB 2     // min resident 256-thread CTAs per SM of the register-resident sweep kernel (register cap)
#endif
#define DRJ_EXCH 4      // doubles per thread exchanged per shared-memory pass during a rung swap

#ifndef DRJACOBY_B200_H  // the host side takes these from include/drjacoby_b200.h
enum {
  DRJ_OK = 0,
  DRJ_ERR_INIT_NEGINF = 1,   // Particle.h:75-78
  DRJ_ERR_LOGLIKE_BAD = 2,   // Particle.h:121-124
  DRJ_ERR_LOGPRIOR_BAD = 3,  // Particle.h:125-128
  DRJ_ERR_TRANS_TYPE = 4     // Particle.cpp:71,95,120
};
#endif

struct DrjKernelArgs {
  int chains;          // chains in this shard
  int rungs;
  int chains_per_cta;
  int n_iter;          // iterations in this launch
  int iter0;           // global update index t (1-based) of the first iteration of this launch
  int coupling_on;
  int save_hot_draws;
  int store_pt;        // ring holds loglike/logprior of all rungs (1) or of the cold rung only (0)
  int rng_mode;        // 0 Philox, 1 replay
  int d_free;
  int use_tma;
  int warp_local;      // rungs divide 32 and nothing in the loop needs a CTA barrier: chains sync per warp
  int team_ctas;       // TEAM kernels: CTAs per cluster (one cluster per chain), 1..16
  int team_stages;     // TEAM kernels: depth of the TMA tile pipeline (tiles resident per CTA), 2..DRJ_TEAM_MAXSTAGES
  int grid_slots;      // GRID kernels: data slots per particle in a CTA (threads slot * P + particle), fixed by the WHOLE run's size
  int grid_stages;     // GRID kernels: depth of the TMA chunk pipeline, 2..DRJ_GRID_MAXSTAGES
  int draws;           // one-thread-per-particle kernels: the state-independent part of every update of this launch (the standard
                       // normal draw and log u of the accept test, log u of every coupling pair) was computed by
                       // drj_draw_kernel into draw_z / draw_lu / draw_clu (runs too small to fill the machine)
  const double* draw_z;    // [n_iter + 1][d_free][P]  (one padded row: the loads run one update ahead)
  const double* draw_lu;   // same layout
  const double* draw_clu;  // [n_iter][chains][rungs - 1]
  int couple_par;      // one-thread-per-particle kernels, rungs <= 32, small runs: the swap decisions of a coupling pass are taken for
                       // every possible carried replica at once, the hot-to-cold walk is integer arithmetic on those masks
  int blocks_trivial;  // a one-block model whose every free parameter lists exactly that block: no block lists are read
  unsigned lo_is_log;  // bit i: parameter i (< 32) has a lower bound of exactly 0, so its cached Jacobian logarithm
                       // log(theta_i - min_i) IS log(theta_i) (models with HINTS reuse it instead of taking the logarithm again)
  int grid_tile;       // GRID kernels: 2..8 particles in the WHOLE run: every thread is a data slot and accumulates for all of them
  long long chain_offset;     // global id of chain 0 (Philox counter)
  unsigned long long seed;
  double target_acceptance;
  long long P;                // chains * rungs
  // per-parameter metadata, device arrays
  const double* theta_min;    // [D]
  const double* theta_max;    // [D]
  const int* trans_type;      // [D]
  const int* free_slot;       // [D] index among free parameters, -1 = skip_param
  const int* block_ptr;       // [D+1]
  const int* block_idx;       // 1-based
  const double* beta_raised;  // [R]
  const double* theta_init;   // [chains][D]
  DrjList data;
  DrjList misc;
  // replay uniforms [chains][T][U]
  const double* replay;
  long long replay_T;
  long long replay_U;
  // particle state, SoA [field][P]
  double* theta;      // [D][P]
  double* phi;        // [D][P]
  double* logbw;      // [D][P]  log of the proposal bandwidth (Particle::bw)
  int* bw_index;      // [D][P]
  double* adj_lo;     // [D][P]  log(theta - min) of the current state
  double* adj_hi;     // [D][P]  log(max - theta) of the current state
  double* ll_block;   // [B][P]
  double* ll;         // [P]
  double* lp;         // [P]
  int* accept;        // [P]
  int* mc_accept;     // [chains][R-1]
  // sample rings of this launch
  double* ring_ll;    // [n_iter][P] or [n_iter][chains]
  double* ring_lp;
  double* ring_theta; // [n_iter][D][P] or [n_iter][D][chains]
  // initial block likelihoods / prior of every chain (all rungs of a chain start from the same theta)
  double* init_llb;   // [B][chains]
  double* init_lp;    // [chains]
  // GRID kernels: per-virtual-rank partial sums [2][DRJ_GRID_VRANKS][P] and the grid barrier counter (0 at launch)
  double* grid_part;
  unsigned int* grid_bar;
  // error slot
  unsigned long long* err_key;  // packed (iteration-in-launch, chain, rung, param), atomicMin
  int* err_code;                // [P] code of the particle's first error in this launch
  double* err_theta;            // [D][P] offending theta'
};

// ------------------------------------------------------------------ TMA bulk copy + mbarrier
__device__ __forceinline__ unsigned drj_smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void drj_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(drj_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void drj_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void drj_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(drj_smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void drj_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "DRJ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DRJ_DONE;\n"
      "bra DRJ_WAIT;\n"
      "DRJ_DONE:\n"
      "}\n" ::"r"(drj_smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void drj_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(drj_smem_addr(bar)) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on `bar`
__device__ __forceinline__ void drj_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                             unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   drj_smem_addr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(drj_smem_addr(bar))
               : "memory");
}

// 128-bit shared-memory load whose position in the instruction stream the compiler must keep
__device__ __forceinline__ double2 drj_lds2(unsigned smem_addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_addr));
  return v;
}

// ------------------------------------------------------------------ thread-block cluster helpers (TEAM kernels)
__device__ __forceinline__ unsigned drj_cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned drj_cluster_id_x() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
// hardware barrier over every thread of the cluster, release/acquire at cluster scope
__device__ __forceinline__ void drj_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the shared memory of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ void drj_st_cluster(double* local_smem_ptr, unsigned rank, double v) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(drj_smem_addr(local_smem_ptr)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

// ------------------------------------------------------------------ model traits helpers
// A model is a struct with
//   static constexpr int D, NBLOCK;  static constexpr bool DATUM_FORM;
//   __device__ static double loglike(const double* theta, const DrjList& data, const DrjList& misc, int block);
//   __device__ static double logprior(const double* theta, const DrjList& misc);
// and, when DATUM_FORM (likelihood of `block` = finish(sum over data item datum_item(block))):
//   struct Consts;  static bool prepare(theta, misc, block, Consts&)   (false = use loglike())
//   static double datum(const Consts&, double x, double acc);  static double finish(const Consts&, double acc, long long n);
//   static int datum_item(int block);

template <class M, bool DF = M::DATUM_FORM>
struct DrjTiles {  // no tiles for generic models
  template <int MODE>
  __device__ __forceinline__ void init(const DrjKernelArgs&) {}
};

// The tile buffers and their mbarriers are static shared memory owned by these two accessors, so
// that every user (kernel body, the noinline sweep routine) sees a pointer the compiler KNOWS is
// shared memory (a pointer carried through a struct degrades to generic LD.E loads).
// The buffer itself is DYNAMIC shared memory sized by the host (drj_tile_smem_bytes): the whole data
// item + slack when it stays resident (n <= 2 tiles), two tiles + slack when it is streamed -- a small
// data set must not cost 32 KB per CTA and cap the number of resident CTAs.
extern __shared__ __align__(128) double drj_dyn_smem[];
template <class M>
__device__ __forceinline__ double* drj_tile_buf() {
  return drj_dyn_smem;
}
#define DRJ_TEAM_TILE 8192     // TEAM kernels: doubles per data tile (64 KB); the per-tile hand-shake (mbarrier wait,
                               // CTA barrier, refill) is a ~700-cycle dependent chain, 16 KB tiles left it dominant
#define DRJ_TEAM_MAXSTAGES 12  // TEAM kernels: at most this many tiles in flight per CTA
#define DRJ_TEAM_VRANKS 16     // TEAM kernels: canonical number of partial sums a data item is split into
#define DRJ_TEAM_MAXR 64       // TEAM kernels: at most this many rungs (sizes the all-gather buffer: 2 x 16 x MAXR doubles)
#ifndef DRJ_TEAM_NT
#define DRJ_TEAM_NT 256        // TEAM kernels: threads per CTA (512, at a 128-register cap, measured 5-10 % slower on every shape)
#endif
// GRID kernels (few particles x data far larger than L2: the HBM-bound regime)
#define DRJ_GRID_VRANKS 1184   // canonical number of partial sums a data item is split into: 8 per SM of a 148-SM B200
#ifndef DRJ_GRID_CHUNK
#define DRJ_GRID_CHUNK 8192    // doubles per TMA chunk (64 KB).  Measured at n = 1e8 (profiles/README.md, round 2): the consumer side
                               // pays ~800 cycles per chunk whatever its size (full/empty hand-shake, LDS latency at loop entry with
                               // 2 warps per scheduler), so 16 KB chunks ran at 0.69 (1 x 1) / 0.44 (5 x 1) of the HBM peak, 64 KB
                               // chunks at 0.98 / 0.80; the number of stages made no difference
#endif
#define DRJ_GRID_MAXSTAGES 12  // at most this many chunks in flight per CTA
#define DRJ_GRID_MAXP 256      // at most this many particles (chains x rungs): every CTA holds a replica of each
// kernel families: one thread per particle / one cluster per chain / the whole grid per run
#define DRJ_MODE_SOLO 0
#define DRJ_MODE_TEAM 1
#define DRJ_MODE_GRID 2
template <class M>
__device__ __forceinline__ unsigned long long* drj_tile_bars() {
  __shared__ __align__(8) unsigned long long s_full[DRJ_TEAM_MAXSTAGES];
  return s_full;
}
// GRID: "stage consumed" barriers (one arrival per warp), so that only the producer thread ever waits for the slowest warp
template <class M>
__device__ __forceinline__ unsigned long long* drj_tile_empty_bars() {
  __shared__ __align__(8) unsigned long long s_empty[DRJ_GRID_MAXSTAGES];
  return s_empty;
}
// TEAM kernels: per-thread partials of the CTA reduction, and the all-gathered partial sums
// [parity][virtual rank][rung] every CTA of the cluster receives from its peers
template <class M>
__device__ __forceinline__ double* drj_team_red() {
  __shared__ double s_red[DRJ_TEAM_NT];
  return s_red;
}
template <class M>
__device__ __forceinline__ double* drj_team_part() {
  __shared__ double s_part[2 * DRJ_TEAM_VRANKS * DRJ_TEAM_MAXR];
  return s_part;
}

template <class M>
struct DrjTiles<M, true> {
  unsigned ph0, ph1;           // phase parity of each stage
  unsigned team_ph;            // TEAM / GRID: phase parity bit of every stage's "full" barrier (consumer side)
  unsigned team_pb;            // TEAM / GRID: which half of the partial-sum buffer the next sweep uses
  unsigned grid_eph;           // GRID, producer thread: phase parity bit of every stage's "empty" barrier
  unsigned grid_epoch;         // GRID: grid barriers passed so far in this launch
  int grid_cs, grid_ps;        // GRID: next stage to consume / to fill
  int grid_pn;                 // GRID, producer thread: chunks requested so far in this launch
  bool resident;               // whole item lives in the tile buffer for the life of the kernel
  bool tma;
  template <int MODE>
  __device__ __forceinline__ void init(const DrjKernelArgs& a) {
    constexpr bool TEAM = MODE != DRJ_MODE_SOLO;  // (both cluster and grid kernels always stream)
    double* s_buf = drj_tile_buf<M>();
    unsigned long long* full = drj_tile_bars<M>();
    ph0 = ph1 = 0;
    team_ph = 0; team_pb = 0;
    grid_eph = 0; grid_epoch = 0; grid_cs = 0; grid_ps = 0; grid_pn = 0;
    tma = a.use_tma != 0;
    // single-item models keep the data resident when it fits the two tiles
    const int item = M::datum_item(1);
    const long long n = a.data.len(item);
    resident = !TEAM && (M::NBLOCK == 1) && (n <= 2 * DRJ_TILE);
    if (threadIdx.x == 0) {
      const int nbar = MODE == DRJ_MODE_GRID ? a.grid_stages : (MODE == DRJ_MODE_TEAM ? a.team_stages : 2);
      for (int q = 0; q < nbar; ++q) drj_mbar_init(&full[q], 1);
      if (MODE == DRJ_MODE_GRID) {
        unsigned long long* empty = drj_tile_empty_bars<M>();
        for (int q = 0; q < nbar; ++q) drj_mbar_init(&empty[q], blockDim.x >> 5);
      }
      drj_mbar_fence_init();
    }
    if (resident) {
      const double* g = a.data.item(item);
      for (long long j = threadIdx.x; j < n; j += blockDim.x) s_buf[j] = g[j];
    }
    __syncthreads();
  }
};

// DRJ_NACC interleaved partial sums: datum j -> accumulator (j mod 8), combined pairwise at the end.
// One dependent DFMA chain per thread cannot keep the FP64 pipe busy (profiles/README.md, round 1:
// 67 % pipe utilisation, "wait" stalls on the accumulate DFMA); tools/fp64_pattern_probe.cu shows the
// pipe needs >= 4-8 independent chains per thread at 16 warps/SM.  The assignment is by GLOBAL datum
// index (tiles start at multiples of 8), so the result does not depend on tile size or launch
// geometry.  It is a reassociation of the reference's sequential `ret +=` loop
// (inst/extdata/example_loglike_logprior.cpp:15-18): same terms, error O(ulp), inside the 1e-10 gate.
#define DRJ_NACC 8
struct DrjAcc {
  double a[DRJ_NACC];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int q = 0; q < DRJ_NACC; ++q) a[q] = 0.0;
  }
  __device__ __forceinline__ double total() const {
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
};

// FULL = the span is a whole tile: compile-time trip count, which lets ptxas keep the loads of the
// next group(s) far ahead of their use (the same source with a run-time count is scheduled with the
// LDS only ~20 cycles ahead and stalls on it, profiles/README.md).
template <class M, bool FULL>
__device__ __forceinline__ void drj_accum_span(const typename M::Consts& k, const double* __restrict__ buf, int cnt,
                                               DrjAcc& A) {
  const int n8 = FULL ? (DRJ_TILE / 8) : (cnt >> 3);
  if (n8 > 0) {
    // software pipeline: the four broadcast LDS.128 of group j+1 are issued BEFORE the 16 DFMA of
    // group j.  The loads are volatile inline PTX: with plain C++ loads ptxas sinks them to ~20
    // cycles before their use and the warp stalls on them (profiles/README.md); pinned like this the
    // pattern runs at 97 % of the DFMA peak in tools/fp64_pattern_probe3.cu.  The last trip
    // prefetches 8 doubles past the span; the tile buffer carries 8 doubles of slack.
    const unsigned base = drj_smem_addr(buf);
    double2 c0 = drj_lds2(base), c1 = drj_lds2(base + 16), c2 = drj_lds2(base + 32), c3 = drj_lds2(base + 48);
#pragma unroll 4
    for (int j = 0; j < n8; ++j) {
      const unsigned ad = base + 64u * (unsigned)j + 64u;
      const double2 n0 = drj_lds2(ad), n1 = drj_lds2(ad + 16), n2 = drj_lds2(ad + 32), n3 = drj_lds2(ad + 48);
      A.a[0] = M::datum(k, c0.x, A.a[0]);
      A.a[1] = M::datum(k, c0.y, A.a[1]);
      A.a[2] = M::datum(k, c1.x, A.a[2]);
      A.a[3] = M::datum(k, c1.y, A.a[3]);
      A.a[4] = M::datum(k, c2.x, A.a[4]);
      A.a[5] = M::datum(k, c2.y, A.a[5]);
      A.a[6] = M::datum(k, c3.x, A.a[6]);
      A.a[7] = M::datum(k, c3.y, A.a[7]);
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
  }
  if (!FULL) {
    const int j0 = n8 << 3;
#pragma unroll
    for (int q = 0; q < DRJ_NACC - 1; ++q)
      if (j0 + q < cnt) A.a[q] = M::datum(k, buf[j0 + q], A.a[q]);
  }
}

// CTA-wide likelihood sweep: every thread evaluates ITS OWN theta against the SAME data item.
// Must be called by all threads of the CTA (uniform control flow).
template <class M>
__device__ __forceinline__ double drj_sweep_item_inl(DrjTiles<M, true>& T, const typename M::Consts& k,
                                                     const double* __restrict__ g, long long n) {
  double* const buf = drj_tile_buf<M>();
  unsigned long long* const full = drj_tile_bars<M>();
  DrjAcc A;
  A.zero();
  if (T.resident) {
    drj_accum_span<M, false>(k, buf, (int)n, A);
    return A.total();
  }
  const long long ntiles = (n + DRJ_TILE - 1) / DRJ_TILE;
  const unsigned tile_bytes = DRJ_TILE * sizeof(double);
  if (T.tma) {
    // prologue: TMA engine fills both stages (the buffer is padded to whole tiles by the host)
    if (threadIdx.x == 0 && ntiles > 0) {
      drj_mbar_expect_tx(&full[0], tile_bytes);
      drj_bulk_g2s(buf, g, tile_bytes, &full[0]);
      if (ntiles > 1) {
        drj_mbar_expect_tx(&full[1], tile_bytes);
        drj_bulk_g2s(buf + DRJ_TILE, g + DRJ_TILE, tile_bytes, &full[1]);
      }
    }
    for (long long kt = 0; kt < ntiles; ++kt) {
      const int s = (int)(kt & 1);
      if (s == 0) { drj_mbar_wait(&full[0], T.ph0); T.ph0 ^= 1u; }
      else        { drj_mbar_wait(&full[1], T.ph1); T.ph1 ^= 1u; }
      const long long rem = n - kt * DRJ_TILE;
      const int cnt = rem < DRJ_TILE ? (int)rem : DRJ_TILE;
      if (cnt == DRJ_TILE) drj_accum_span<M, true>(k, buf + s * DRJ_TILE, cnt, A);
      else drj_accum_span<M, false>(k, buf + s * DRJ_TILE, cnt, A);
      __syncthreads();  // everyone is done with stage s before the engine refills it
      if (threadIdx.x == 0 && kt + 2 < ntiles) {
        drj_mbar_expect_tx(&full[s], tile_bytes);
        drj_bulk_g2s(buf + s * DRJ_TILE, g + (kt + 2) * DRJ_TILE, tile_bytes, &full[s]);
      }
    }
  } else {
    // plain cooperative staging (debug / comparison path, selected with DRJ_NO_TMA=1)
    for (long long kt = 0; kt < ntiles; ++kt) {
      const long long rem = n - kt * DRJ_TILE;
      const int cnt = rem < DRJ_TILE ? (int)rem : DRJ_TILE;
      __syncthreads();
      for (int j = threadIdx.x; j < cnt; j += blockDim.x) buf[j] = g[kt * DRJ_TILE + j];
      __syncthreads();
      drj_accum_span<M, false>(k, buf, cnt, A);
    }
  }
  return A.total();
}

// out-of-line copy for the register-resident (non-STREAM) kernel, whose live particle state would
// otherwise be dragged through the sweep's register allocation
template <class M>
__device__ __noinline__ double drj_sweep_item(DrjTiles<M, true>& T, const typename M::Consts& k,
                                              const double* __restrict__ g, long long n) {
  return drj_sweep_item_inl<M>(T, k, g, n);
}

// ------------------------------------------------------------------ TEAM sweep (few chains x long data)
// One thread per (chain, rung) cannot use the machine when a run has few chains: the reference's usual
// shape is 1-10 chains over a long data vector.  The TEAM kernels give every chain a whole thread-block
// CLUSTER: all threads of all CTAs of the cluster hold a replica of the chain's particle states (thread
// tid = slot * R + rung replicates rung `rung`; the update arithmetic is deterministic, so the replicas
// stay bit-identical without communication), and only the data sweep is divided: the tiles of the item
// (DRJ_TEAM_TILE = 8192 doubles) are dealt round-robin to DRJ_TEAM_VRANKS (16) "virtual ranks", CTA c of a K-CTA cluster streams the
// virtual ranks c, c+K, ... through an NS-deep TMA pipeline, inside a tile slot s handles the 16-byte
// pairs s, s+S, ... (S = 256 / R slots per rung) -- lanes of one rung-group read consecutive pairs,
// lanes of one slot broadcast -- and the per-thread partial sums are combined by a fixed tree over the
// slots (shared memory), written to EVERY CTA of the cluster through distributed shared memory, and
// after one hardware cluster barrier summed in virtual-rank order by every thread.  The association
// depends only on (R, n): the result is bitwise independent of the cluster size K and of how chains are
// sharded over GPUs.  HBM / L2 traffic is one pass over the item per chain per sweep.
__device__ __forceinline__ int drj_team_pairs(int cnt, int slot, int S) {
  const int np = cnt >> 1;  // whole 16-byte pairs in the tile
  return np > slot ? (np - slot + S - 1) / S : 0;
}
template <class M>
__device__ __forceinline__ void drj_team_accum(const typename M::Consts& k, const double* __restrict__ buf, int cnt,
                                               int slot, int S, int mine, DrjAcc& A) {
  // `mine` = drj_team_pairs(cnt, slot, S): the 16-byte pairs slot, slot + S, ... of this thread in the tile.
  // (plain loads: pinning the next group's loads ahead of the arithmetic, as drj_accum_span does, measured
  // 10-35 % SLOWER here; a counted loop the compiler can unroll is what helps -- tools/team_bench.py, round 1)
  const int np = cnt >> 1;
  const double2* __restrict__ pq = reinterpret_cast<const double2*>(buf) + slot;
  const int ngroups = mine >> 2;
#pragma unroll 4
  for (int gi = 0; gi < ngroups; ++gi) {
    const double2 v0 = pq[0], v1 = pq[S], v2 = pq[2 * S], v3 = pq[3 * S];
    pq += 4 * S;
    A.a[0] = M::datum(k, v0.x, A.a[0]);
    A.a[1] = M::datum(k, v0.y, A.a[1]);
    A.a[2] = M::datum(k, v1.x, A.a[2]);
    A.a[3] = M::datum(k, v1.y, A.a[3]);
    A.a[4] = M::datum(k, v2.x, A.a[4]);
    A.a[5] = M::datum(k, v2.y, A.a[5]);
    A.a[6] = M::datum(k, v3.x, A.a[6]);
    A.a[7] = M::datum(k, v3.y, A.a[7]);
  }
  for (int j = 0; j < (mine & 3); ++j) {
    const double2 v = pq[0];
    pq += S;
    A.a[0] = M::datum(k, v.x, A.a[0]);
    A.a[1] = M::datum(k, v.y, A.a[1]);
  }
  if ((cnt & 1) && slot == np % S) A.a[2] = M::datum(k, buf[cnt - 1], A.a[2]);
}

// Must be called by every thread of every CTA of the cluster (uniform control flow).  Inlined: the
// pipeline state (phase bits) stays in registers; the TEAM kernels run one CTA per SM and may use the
// whole register file.
template <class M>
__device__ __forceinline__ double drj_sweep_item_team(DrjTiles<M, true>& T, const DrjKernelArgs& a,
                                                      const typename M::Consts& k, const double* __restrict__ g,
                                                      long long n) {
  double* const buf = drj_tile_buf<M>();
  unsigned long long* const full = drj_tile_bars<M>();
  double* const s_red = drj_team_red<M>();
  double* const s_part = drj_team_part<M>() + T.team_pb * (DRJ_TEAM_VRANKS * DRJ_TEAM_MAXR);
  T.team_pb ^= 1u;  // the peers may still be reading the other half (one cluster barrier per sweep)
  const int R = a.rungs, S = a.chains_per_cta, K = a.team_ctas, NS = a.team_stages;
  const int tid = threadIdx.x;
  const int slot = tid / R, r = tid - slot * R;
  const int crank = (int)drj_cluster_ctarank();
  const int ntiles = (int)((n + DRJ_TEAM_TILE - 1) / DRJ_TEAM_TILE);
  const int last_cnt = (int)(n - (long long)(ntiles - 1) * DRJ_TEAM_TILE);  // data points in the last tile

  // producer cursor (thread 0): the next tile to request, in the order this CTA consumes them.  The
  // copy is exactly the tile's data rounded up to 16 bytes (items are padded to even lengths).
  int pv = crank, pt = crank, pstage = 0;
  auto produce = [&]() {
    while (pv < DRJ_TEAM_VRANKS && pt >= ntiles) { pv += K; pt = pv; }
    if (pv >= DRJ_TEAM_VRANKS) return;
    const int c = (pt == ntiles - 1) ? last_cnt : DRJ_TEAM_TILE;
    const unsigned bytes = (unsigned)((c + 1) & ~1) * (unsigned)sizeof(double);
    drj_mbar_expect_tx(&full[pstage], bytes);
    drj_bulk_g2s(buf + pstage * DRJ_TEAM_TILE, g + (long long)pt * DRJ_TEAM_TILE, bytes, &full[pstage]);
    pt += DRJ_TEAM_VRANKS;
    pstage = (pstage + 1 == NS) ? 0 : pstage + 1;
  };
  if (T.tma && tid == 0)
    for (int q = 0; q < NS; ++q) produce();

  int stage = 0;
  unsigned ph = T.team_ph;
  const int mine_full = drj_team_pairs(DRJ_TEAM_TILE, slot, S);  // (an integer division: once per sweep, not per tile)
  for (int v = crank; v < DRJ_TEAM_VRANKS && v < ntiles; v += K) {
    DrjAcc A;
    A.zero();
    for (int kt = v; kt < ntiles; kt += DRJ_TEAM_VRANKS) {
      const int cnt = (kt == ntiles - 1) ? last_cnt : DRJ_TEAM_TILE;
      if (T.tma) {
        drj_mbar_wait(&full[stage], (ph >> stage) & 1u);
        ph ^= 1u << stage;
        if (slot < S)
          drj_team_accum<M>(k, buf + stage * DRJ_TEAM_TILE, cnt, slot, S,
                            cnt == DRJ_TEAM_TILE ? mine_full : drj_team_pairs(cnt, slot, S), A);
        __syncthreads();  // everyone is done with this stage before the engine refills it
        if (tid == 0) produce();
        stage = (stage + 1 == NS) ? 0 : stage + 1;
      } else {
        __syncthreads();
        for (int j = tid; j < cnt; j += blockDim.x) buf[j] = g[(long long)kt * DRJ_TEAM_TILE + j];
        __syncthreads();
        if (slot < S)
          drj_team_accum<M>(k, buf, cnt, slot, S, cnt == DRJ_TEAM_TILE ? mine_full : drj_team_pairs(cnt, slot, S), A);
      }
    }
    // fixed tree over the S slots of every rung
    s_red[tid] = (slot < S) ? A.total() : 0.0;
    __syncthreads();
    for (int m = S; m > 1;) {
      const int half = (m + 1) >> 1;
      if (slot < m - half) s_red[tid] += s_red[tid + half * R];
      __syncthreads();
      m = half;
    }
    if (tid < R) {
      const double part = s_red[tid];
      for (int c = 0; c < K; ++c) drj_st_cluster(&s_part[v * DRJ_TEAM_MAXR + tid], (unsigned)c, part);
    }
  }
  T.team_ph = ph;
  drj_cluster_sync();
  const int nv = ntiles < DRJ_TEAM_VRANKS ? ntiles : DRJ_TEAM_VRANKS;
  double tot = 0.0;
  for (int v = 0; v < nv; ++v) tot += s_part[v * DRJ_TEAM_MAXR + r];
  return tot;
}

// ------------------------------------------------------------------ GRID sweep (few particles x data >> L2)
// The HBM-bound regime (SURVEY 7.1-4 "for few chains x huge n, switch to split-n"): with a handful of particles
// over a data item far larger than L2 the likelihood sum costs 8 bytes of HBM traffic per 2 P FP64 instructions,
// and the only way to the HBM roofline is to read the item ONCE per sweep with every SM.  The GRID kernels run
// the whole grid as one team: every CTA holds a replica of EVERY particle (thread tid < P replicates particle
// tid = chain * R + rung; the update arithmetic is deterministic, so the replicas of all CTAs stay bit-identical
// without communication, as in the TEAM kernels), and only the data sweep is divided:
//   * the item is cut into chunks of DRJ_GRID_CHUNK doubles and the chunks, in order, into DRJ_GRID_VRANKS
//     contiguous "virtual ranks" (fewer when the item has fewer chunks); CTA b of a G-CTA grid owns the
//     contiguous virtual ranks [b V / G, (b + 1) V / G), i.e. one contiguous 1/G of the item;
//   * the CTA streams its chunks through an NS-deep TMA pipeline (cp.async.bulk + one "full" mbarrier per stage;
//     a stage is handed back through an "empty" mbarrier with one arrival per warp, and the producer thread
//     refills the stage released one chunk earlier, so no warp but the producer's ever waits for another);
//   * inside a chunk the thread (slot s, particle q) = tid s * P + q takes the 16-byte pairs s, s + S, ... and
//     accumulates them for ITS particle's proposal (Consts published through shared memory): one pass over the
//     data serves all P particles;
//   * at the end of a virtual rank the S per-slot sums of every particle are combined in a fixed two-level order
//     and written to global memory; ONE grid-wide barrier per sweep (co-residency is guaranteed by the
//     cooperative launch); then every CTA sums the <= 1184 partials of every particle in the same fixed order
//     (lane l takes v = l, l + 32, ... sequentially, xor-shuffle tree).
// The association depends only on (n, S): results are bitwise independent of the grid size, the pipeline depth
// and the launch split; S is derived from the WHOLE run's particle count (drj_config.layout_chains), so a shard
// of a run reproduces its slice bit for bit.  HBM traffic: 8 n bytes per sweep for all particles together.
__device__ __forceinline__ void drj_grid_barrier(unsigned int* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // the CTA's partial sums (ordered before by the CTA barrier) become visible grid-wide
    atomicAdd(bar, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

// first chunk of virtual rank v when `nchunks` chunks are dealt to `nv` ranks in contiguous, near-equal runs
__device__ __forceinline__ long long drj_grid_first_chunk(long long v, long long nchunks, long long nv) {
  return (v * nchunks) / nv;
}

// 2 to 8 particles (the reference's default is five chains): with one thread per (slot, particle) every particle's
// thread loads the datum again from shared memory -- P loads per datum, and at 5 x 1 the sweep was bound by them (0.80 of
// the HBM peak, FP64 pipe 36 %, short_scoreboard 1.8).  Here every thread is a slot, loads a 16-byte pair ONCE and
// accumulates it for all P particles (their Consts and two partial sums each in registers).  Pair p of the item goes to
// thread p mod 256 (chunks hold a multiple of 256 pairs), its two data to that thread's two sums: the association
// does not depend on the chunking, the grid size or the launch split.
#define DRJ_GRID_TILE_MAXP 8
template <class M, int NP>
__device__ __forceinline__ void drj_grid_accum_tile_n(const typename M::Consts (&kq)[DRJ_GRID_TILE_MAXP],
                                                      const double* __restrict__ buf, int cnt, int tid, int nthreads,
                                                      double (&acc)[DRJ_GRID_TILE_MAXP][2]) {
  const int np = cnt >> 1;
  const double2* __restrict__ pq = reinterpret_cast<const double2*>(buf);
#pragma unroll 4
  for (int pi = tid; pi < np; pi += nthreads) {
    const double2 v = pq[pi];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      acc[q][0] = M::datum(kq[q], v.x, acc[q][0]);
      acc[q][1] = M::datum(kq[q], v.y, acc[q][1]);
    }
  }
  if ((cnt & 1) && tid == np % nthreads) {
#pragma unroll
    for (int q = 0; q < NP; ++q) acc[q][0] = M::datum(kq[q], buf[cnt - 1], acc[q][0]);
  }
}
// (the particle count as a compile-time constant: predicated-off instructions of absent particles would still issue)
template <class M>
__device__ __forceinline__ void drj_grid_accum_tile(const typename M::Consts (&kq)[DRJ_GRID_TILE_MAXP], int P,
                                                    const double* __restrict__ buf, int cnt, int tid, int nthreads,
                                                    double (&acc)[DRJ_GRID_TILE_MAXP][2]) {
  switch (P) {
    case 1: drj_grid_accum_tile_n<M, 1>(kq, buf, cnt, tid, nthreads, acc); break;
    case 2: drj_grid_accum_tile_n<M, 2>(kq, buf, cnt, tid, nthreads, acc); break;
    case 3: drj_grid_accum_tile_n<M, 3>(kq, buf, cnt, tid, nthreads, acc); break;
    case 4: drj_grid_accum_tile_n<M, 4>(kq, buf, cnt, tid, nthreads, acc); break;
    case 5: drj_grid_accum_tile_n<M, 5>(kq, buf, cnt, tid, nthreads, acc); break;
    case 6: drj_grid_accum_tile_n<M, 6>(kq, buf, cnt, tid, nthreads, acc); break;
    case 7: drj_grid_accum_tile_n<M, 7>(kq, buf, cnt, tid, nthreads, acc); break;
    default: drj_grid_accum_tile_n<M, 8>(kq, buf, cnt, tid, nthreads, acc); break;
  }
}

// Must be called by every thread of every CTA of the grid (uniform control flow).  `own` = index of the particle
// whose total this thread wants back (its own particle; shadow threads replicate a particle of chain 0).
template <class M>
__device__ __forceinline__ double drj_sweep_item_grid(DrjTiles<M, true>& T, const DrjKernelArgs& a,
                                                      const typename M::Consts& kmine, int own,
                                                      const double* __restrict__ g, long long n) {
  __shared__ typename M::Consts s_consts[DRJ_GRID_MAXP];
  __shared__ double s_l1[DRJ_NT];
  double* const buf = drj_tile_buf<M>();
  unsigned long long* const full = drj_tile_bars<M>();
  unsigned long long* const empty = drj_tile_empty_bars<M>();
  double* const s_red = drj_team_red<M>();
  const int P = (int)a.P, S = a.grid_slots, NS = a.grid_stages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int slot = tid / P, q = tid - slot * P;
  const bool worker = slot < S;
  double* const part = a.grid_part + (long long)T.team_pb * DRJ_GRID_VRANKS * P;
  T.team_pb ^= 1u;  // slower CTAs may still be reading the other half (one grid barrier per sweep)

  // every particle's proposal constants, published by the replica in slot 0
  if (tid < P) s_consts[tid] = kmine;
  __syncthreads();
  typename M::Consts k;
  if (worker) k = s_consts[q];
  const bool tile = a.grid_tile != 0 && P <= DRJ_GRID_TILE_MAXP;
  typename M::Consts kq[DRJ_GRID_TILE_MAXP];
  if (tile) {
#pragma unroll
    for (int qq = 0; qq < DRJ_GRID_TILE_MAXP; ++qq) kq[qq] = s_consts[qq < P ? qq : 0];
  }
  __shared__ double s_wsum[DRJ_GRID_TILE_MAXP][DRJ_NT / 32];

  const long long nchunks = (n + DRJ_GRID_CHUNK - 1) / DRJ_GRID_CHUNK;
  const long long nv = nchunks < DRJ_GRID_VRANKS ? nchunks : DRJ_GRID_VRANKS;
  const long long G = gridDim.x, b = blockIdx.x;
  const long long v_lo = (b * nv) / G, v_hi = ((b + 1) * nv) / G;
  const long long c_lo = nv ? drj_grid_first_chunk(v_lo, nchunks, nv) : 0;
  const long long c_hi = nv ? drj_grid_first_chunk(v_hi, nchunks, nv) : 0;
  const int last_cnt = (int)(n - (nchunks - 1) * DRJ_GRID_CHUNK);  // data points in the item's last chunk

  // producer (thread 0): chunk `pc` goes into stage T.grid_ps once the warps have handed that stage back
  long long pc = c_lo;
  auto produce = [&]() {
    if (pc >= c_hi) return;
    const int ps = T.grid_ps;
    if (T.grid_pn >= NS) {
      drj_mbar_wait(&empty[ps], (T.grid_eph >> ps) & 1u);
      T.grid_eph ^= 1u << ps;
    }
    const int c = (pc == nchunks - 1) ? last_cnt : DRJ_GRID_CHUNK;
    const unsigned bytes = (unsigned)((c + 1) & ~1) * (unsigned)sizeof(double);
    drj_mbar_expect_tx(&full[ps], bytes);
    drj_bulk_g2s(buf + ps * DRJ_GRID_CHUNK, g + pc * DRJ_GRID_CHUNK, bytes, &full[ps]);
    ++pc;
    ++T.grid_pn;
    T.grid_ps = (ps + 1 == NS) ? 0 : ps + 1;
  };
  if (T.tma && tid == 0)
    for (int j = 0; j < NS - 1; ++j) produce();

  const int mine_full = worker ? drj_team_pairs(DRJ_GRID_CHUNK, slot, S) : 0;
  // level-1 groups of the per-rank reduction: J groups of slots j, j + J, ... per particle
  const int J = S < 16 ? S : 16;
  long long c = c_lo;
  for (long long v = v_lo; v < v_hi; ++v) {
    const long long c_end = drj_grid_first_chunk(v + 1, nchunks, nv);
    DrjAcc A;
    A.zero();
    double acc[DRJ_GRID_TILE_MAXP][2];
#pragma unroll
    for (int qq = 0; qq < DRJ_GRID_TILE_MAXP; ++qq) { acc[qq][0] = 0.0; acc[qq][1] = 0.0; }
    for (; c < c_end; ++c) {
      const int cnt = (c == nchunks - 1) ? last_cnt : DRJ_GRID_CHUNK;
      if (T.tma) {
        if (tid == 0) produce();  // refills the stage of chunk c - 1
        const int cs = T.grid_cs;
        drj_mbar_wait(&full[cs], (T.team_ph >> cs) & 1u);
        T.team_ph ^= 1u << cs;
        if (tile) drj_grid_accum_tile<M>(kq, P, buf + cs * DRJ_GRID_CHUNK, cnt, tid, (int)blockDim.x, acc);
        else if (worker)
          drj_team_accum<M>(k, buf + cs * DRJ_GRID_CHUNK, cnt, slot, S,
                            cnt == DRJ_GRID_CHUNK ? mine_full : drj_team_pairs(cnt, slot, S), A);
        __syncwarp();
        if (lane == 0) drj_mbar_arrive(&empty[cs]);
        T.grid_cs = (cs + 1 == NS) ? 0 : cs + 1;
      } else {
        __syncthreads();
        for (int j = tid; j < cnt; j += blockDim.x) buf[j] = g[c * DRJ_GRID_CHUNK + j];
        __syncthreads();
        if (tile) drj_grid_accum_tile<M>(kq, P, buf, cnt, tid, (int)blockDim.x, acc);
        else if (worker) drj_team_accum<M>(k, buf, cnt, slot, S, cnt == DRJ_GRID_CHUNK ? mine_full : drj_team_pairs(cnt, slot, S), A);
      }
    }
    if (tile) {
      // per particle: the 32 slot sums of a warp by an xor-shuffle tree, the warps' sums in order
#pragma unroll
      for (int qq = 0; qq < DRJ_GRID_TILE_MAXP; ++qq)
        if (qq < P) {
          double t = acc[qq][0] + acc[qq][1];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) s_wsum[qq][tid >> 5] = t;
        }
      __syncthreads();
      if (tid < P) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_wsum[tid][w];
        part[v * P + tid] = t;
      }
      __syncthreads();  // (s_wsum is rewritten by the next virtual rank)
      continue;
    }
    // the S slot sums of every particle, fixed two-level order: group j takes slots j, j + J, ... in sequence,
    // then the J group sums in sequence
    s_red[tid] = worker ? A.total() : 0.0;  // (the previous rank's reads of s_red / s_l1 are two CTA barriers back)
    __syncthreads();
    if (tid < P * J) {
      const int qq = tid % P, j = tid / P;
      double t = 0.0;
      for (int s2 = j; s2 < S; s2 += J) t += s_red[s2 * P + qq];
      s_l1[tid] = t;
    }
    __syncthreads();
    if (tid < P) {
      double t = 0.0;
      for (int j = 0; j < J; ++j) t += s_l1[j * P + tid];
      part[v * P + tid] = t;
    }
  }
  ++T.grid_epoch;
  drj_grid_barrier(a.grid_bar, T.grid_epoch * (unsigned)G);
  // totals: warp w takes the particles w, w + nwarps, ...; lane l the partials l, l + 32, ... of the particle
  {
    const int nwarps = blockDim.x >> 5;
    for (int qq = tid >> 5; qq < P; qq += nwarps) {
      double t = 0.0;
#pragma unroll 4
      for (long long v = lane; v < nv; v += 32) t += __ldcg(&part[v * P + qq]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) s_l1[qq] = t;
    }
  }
  __syncthreads();
  const double tot = s_l1[own];
  __syncthreads();  // (s_l1 is reused by the next sweep)
  return tot;
}

// Optional model hook (HINTS): the sampler already holds log(theta_i - min_i) of every parameter with a lower bound (the
// Jacobian term of the reparameterisation, Particle.cpp:100-118); where min_i is exactly 0 that IS log(theta_i), the value
// a scale parameter's likelihood (log sd) and a log-normal or gamma prior (log x) compute again -- the n = 100 normal
// model took log(sigma) three times per sigma update.  A model with `static constexpr bool HINTS = true` provides
//   prepare_h(theta, misc, block, k, loglo, mask)   [per-datum models]   and   logprior_h(theta, misc, loglo, mask)
// and may use loglo[i] for log(theta[i]) when bit i of mask is set: the same function of the same argument, so the
// same bits.
template <class M, class = void>
struct DrjHints {
  template <class K>  // (K = M::Consts; a template so that models without a per-datum form can instantiate the class)
  __device__ static __forceinline__ bool prepare(const double* th, const DrjList& misc, int block, K& k,
                                                 const double*, unsigned) { return M::prepare(th, misc, block, k); }
  __device__ static __forceinline__ double logprior(const double* th, const DrjList& misc, const double*, unsigned) {
    return M::logprior(th, misc);
  }
};
template <class M>
struct DrjHints<M, decltype((void)M::HINTS)> {
  template <class K>
  __device__ static __forceinline__ bool prepare(const double* th, const DrjList& misc, int block, K& k,
                                                 const double* loglo, unsigned mask) { return M::prepare_h(th, misc, block, k, loglo, mask); }
  __device__ static __forceinline__ double logprior(const double* th, const DrjList& misc, const double* loglo, unsigned mask) {
    return M::logprior_h(th, misc, loglo, mask);
  }
};

// Optional model hook (CTA_INIT): `static constexpr bool CTA_INIT = true` + `cta_init(data, misc)`, called by every
// thread of every thread block before the block evaluates the model for the first time (followed by a block barrier):
// the place to put quantities that depend on the data alone into shared memory the model owns.
template <class M, class = void>
struct DrjCtaInit {
  __device__ static __forceinline__ void run(const DrjKernelArgs&) {}
};
template <class M>
struct DrjCtaInit<M, decltype((void)M::CTA_INIT)> {
  __device__ static __forceinline__ void run(const DrjKernelArgs& a) {
    M::cta_init(a.data, a.misc);
    __syncthreads();
  }
};

template <class M, int MODE>
__device__ __forceinline__ double drj_eval_loglike(DrjTiles<M, true>& T, const DrjKernelArgs& a,
                                                   const double* theta, int block, int own, const double* loglo = nullptr) {
  typename M::Consts k;
  const bool fast = DrjHints<M>::prepare(theta, a.misc, block, k, loglo, loglo ? a.lo_is_log : 0u);
  const int item = M::datum_item(block);
  const long long n = a.data.len(item);
  double acc;  // uniform: all threads sweep
  if constexpr (MODE == DRJ_MODE_GRID) acc = drj_sweep_item_grid<M>(T, a, k, own, a.data.item(item), n);
  else if constexpr (MODE == DRJ_MODE_TEAM) acc = drj_sweep_item_team<M>(T, a, k, a.data.item(item), n);
  else acc = drj_sweep_item<M>(T, k, a.data.item(item), n);
  if (fast) return M::finish(k, acc, n);
  return M::loglike(theta, a.data, a.misc, block);  // degenerate theta: exact serial evaluation
}

template <class M, int MODE>
__device__ __forceinline__ double drj_eval_loglike(DrjTiles<M, false>&, const DrjKernelArgs& a,
                                                   const double* theta, int block, int, const double* = nullptr) {
  return M::loglike(theta, a.data, a.misc, block);
}

// ------------------------------------------------------------------ error slot
__device__ __forceinline__ void drj_report(const DrjKernelArgs& a, int code, int it, long long chain,
                                           int r, int param, long long p, const double* theta, int D) {
  if (a.err_code[p] != 0) return;  // keep the particle's first error
  a.err_code[p] = code;
  for (int i = 0; i < D; ++i) a.err_theta[(long long)i * a.P + p] = theta[i];
  unsigned long long key = ((unsigned long long)(unsigned)it << 44) | ((unsigned long long)chain << 20) |
                           ((unsigned long long)(unsigned)r << 10) | (unsigned long long)(unsigned)(param + 1);
  atomicMin(a.err_key, key);
}

// ------------------------------------------------------------------ per-thread particle state
template <class M>
struct DrjState {
  double theta[M::D], phi[M::D], logbw[M::D], alo[M::D], ahi[M::D];
  int bwi[M::D];
  double llb[M::NBLOCK], llb_p[M::NBLOCK];
  double ll, lp;
  int acc;
};

// small arrays: full unroll (state stays in registers); large arrays live in local memory anyway, and a
// partial unroll of 8 lets the loads of a copy / sum loop overlap instead of paying the L1/L2 latency one
// by one (K4, d = 50: long_scoreboard was 7.2 stalls per issue, profiles/README.md)
#define DRJ_UNROLL_N(N) ((N) <= 8 ? (N) : 8)

// particle state <-> its SoA slots in HBM (launch boundaries; and around a streamed data sweep)
template <class M>
__device__ __forceinline__ void drj_load_state(const DrjKernelArgs& a, DrjState<M>& s, long long p) {
  constexpr int D = M::D, B = M::NBLOCK;
  constexpr int UD = DRJ_UNROLL_N(D), UB = DRJ_UNROLL_N(B);
  const long long P = a.P;
#pragma unroll UD
  for (int i = 0; i < D; ++i) {
    const long long o = (long long)i * P + p;
    s.theta[i] = a.theta[o]; s.phi[i] = a.phi[o]; s.logbw[i] = a.logbw[o];
    s.alo[i] = a.adj_lo[o]; s.ahi[i] = a.adj_hi[o]; s.bwi[i] = a.bw_index[o];
  }
#pragma unroll UB
  for (int b = 0; b < B; ++b) s.llb[b] = a.ll_block[(long long)b * P + p];
  s.ll = a.ll[p]; s.lp = a.lp[p]; s.acc = a.accept[p];
}

template <class M>
__device__ __forceinline__ void drj_store_state(const DrjKernelArgs& a, const DrjState<M>& s, long long p) {
  constexpr int D = M::D, B = M::NBLOCK;
  constexpr int UD = DRJ_UNROLL_N(D), UB = DRJ_UNROLL_N(B);
  const long long P = a.P;
#pragma unroll UD
  for (int i = 0; i < D; ++i) {
    const long long o = (long long)i * P + p;
    a.theta[o] = s.theta[i]; a.phi[o] = s.phi[i]; a.logbw[o] = s.logbw[i];
    a.adj_lo[o] = s.alo[i]; a.adj_hi[o] = s.ahi[i]; a.bw_index[o] = s.bwi[i];
  }
#pragma unroll UB
  for (int b = 0; b < B; ++b) a.ll_block[(long long)b * P + p] = s.llb[b];
  a.ll[p] = s.ll; a.lp[p] = s.lp; a.accept[p] = s.acc;
}

// ------------------------------------------------------------------ init kernel body
// Particle::init + theta_to_phi + init_like for every (chain, rung); writes state and row 0 of the
// burn-in outputs (main.cpp:128-137).
// Particle::init_like (Particle.h:59-79), ONE thread per chain: every rung of a chain starts from the same
// theta (R/main.R:322-328, one init column per chain), so the initial block likelihoods and the prior are the
// same for all of them -- the reference evaluates them once per rung, here once per chain (identical values:
// same inputs, same arithmetic).  With 32 rungs and n = 1e7 this takes the initial sweep from 470 ms to ~25 ms.
// The particle-level init kernel (drj_init_body) then reads them.
template <class M>
__device__ __forceinline__ void drj_init_chain_body(const DrjKernelArgs& a) {
  constexpr int D = M::D, B = M::NBLOCK;
  constexpr int UB = DRJ_UNROLL_N(B);
  DrjCtaInit<M>::run(a);
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = c < a.chains;
  const long long cc = active ? c : 0;  // inactive threads shadow chain 0 (uniform control flow)
  DrjTiles<M> T;
  T.template init<DRJ_MODE_SOLO>(a);
  double theta[D], llb[B];
  for (int i = 0; i < D; ++i) theta[i] = a.theta_init[cc * D + i];
#pragma unroll UB
  for (int b = 0; b < B; ++b) llb[b] = 0.0;
  // every block listed by every parameter (Particle.h:62-68); a block listed by several parameters is
  // evaluated once: same theta, same value
  unsigned long long seen = 0ULL;
  for (int i = 0; i < D; ++i)
    for (int j = a.block_ptr[i]; j < a.block_ptr[i + 1]; ++j) {
      const int b = a.block_idx[j];
      if (b <= 64) {
        if ((seen >> (b - 1)) & 1ULL) continue;
        seen |= 1ULL << (b - 1);
      }
      llb[B == 1 ? 0 : b - 1] = drj_eval_loglike<M, DRJ_MODE_SOLO>(T, a, theta, b, 0);
    }
  const double lp = M::logprior(theta, a.misc);
  if (active) {
#pragma unroll UB
    for (int b = 0; b < B; ++b) a.init_llb[(long long)b * a.chains + c] = llb[b];
    a.init_lp[c] = lp;
  }
}

// TEAM: one cluster per chain, every thread a replica of rung (tid mod R); the replica in slot 0 of CTA 0
// of the cluster (`active`) is the one that writes to HBM.
// GRID: every CTA holds a replica of every particle (thread tid < P = particle tid); CTA 0's write to HBM.
template <class M, int MODE>
__device__ __forceinline__ void drj_init_body(const DrjKernelArgs& a) {
  DrjCtaInit<M>::run(a);
  constexpr bool TEAM = MODE == DRJ_MODE_TEAM, GRID = MODE == DRJ_MODE_GRID;
  constexpr int D = M::D, B = M::NBLOCK;
  constexpr int UD = DRJ_UNROLL_N(D), UB = DRJ_UNROLL_N(B);
  const int R = a.rungs;
  const int tid = threadIdx.x;
  const int lc = tid / R, r = tid - lc * R;
  const long long chain = TEAM ? (long long)drj_cluster_id_x() : (GRID ? (long long)lc : (long long)blockIdx.x * a.chains_per_cta + lc);
  const bool valid = TEAM || (GRID ? (chain < a.chains) : ((lc < a.chains_per_cta) && (chain < a.chains)));
  const bool active = TEAM ? (lc == 0 && drj_cluster_ctarank() == 0u) : (valid && (!GRID || blockIdx.x == 0));
  const long long cc = valid ? chain : 0;  // threads without a particle shadow chain 0 (uniform control flow)
  const long long p = cc * R + r;
  const long long P = a.P;

  DrjTiles<M> T;
  if constexpr (MODE != DRJ_MODE_SOLO) T.template init<MODE>(a);

  DrjState<M> s;
  int err = 0;
#pragma unroll UD
  for (int i = 0; i < D; ++i) {
    const double th = a.theta_init[cc * D + i];
    const double tmin = a.theta_min[i], tmax = a.theta_max[i];
    s.theta[i] = th;
    s.alo[i] = 0.0;
    s.ahi[i] = 0.0;
    switch (a.trans_type[i]) {
      case 0: s.phi[i] = th; break;
      case 1: s.ahi[i] = log(tmax - th); s.phi[i] = s.ahi[i]; break;
      case 2: s.alo[i] = log(th - tmin); s.phi[i] = s.alo[i]; break;
      case 3: s.alo[i] = log(th - tmin); s.ahi[i] = log(tmax - th); s.phi[i] = s.alo[i] - s.ahi[i]; break;
      default: err = DRJ_ERR_TRANS_TYPE; s.phi[i] = th; break;
    }
    s.logbw[i] = 0.0;  // bw = 1
    s.bwi[i] = 1;
  }
#pragma unroll UB
  for (int b = 0; b < B; ++b) s.llb[b] = 0.0;
  if constexpr (MODE != DRJ_MODE_SOLO) {
    // init_like: every block listed by every parameter (duplicates re-evaluated), Particle.h:62-68
    // (a block listed by several parameters is evaluated once: same theta, same value)
    unsigned long long seen = 0ULL;
    for (int i = 0; i < D; ++i)
      for (int j = a.block_ptr[i]; j < a.block_ptr[i + 1]; ++j) {
        const int b = a.block_idx[j];
        if (b <= 64) {
          if ((seen >> (b - 1)) & 1ULL) continue;
          seen |= 1ULL << (b - 1);
        }
        s.llb[B == 1 ? 0 : b - 1] = drj_eval_loglike<M, MODE>(T, a, s.theta, b, (int)p);
      }
    s.lp = M::logprior(s.theta, a.misc);
  } else {
    // evaluated once per chain by drj_init_chain_body (all rungs of a chain start from the same theta)
#pragma unroll UB
    for (int b = 0; b < B; ++b) s.llb[b] = a.init_llb[(long long)b * a.chains + cc];
    s.lp = a.init_lp[cc];
  }
  s.ll = 0.0;
#pragma unroll UB
  for (int b = 0; b < B; ++b) s.ll += s.llb[b];
  if (err == 0 && (s.ll == -DRJ_INF || s.lp == -DRJ_INF)) err = DRJ_ERR_INIT_NEGINF;

  if (active) {
    if (err) drj_report(a, err, 0, chain, r, -1, p, s.theta, D);
#pragma unroll UD
    for (int i = 0; i < D; ++i) {
      const long long o = (long long)i * P + p;
      a.theta[o] = s.theta[i]; a.phi[o] = s.phi[i]; a.logbw[o] = s.logbw[i];
      a.adj_lo[o] = s.alo[i]; a.adj_hi[o] = s.ahi[i]; a.bw_index[o] = s.bwi[i];
    }
#pragma unroll UB
    for (int b = 0; b < B; ++b) a.ll_block[(long long)b * P + p] = s.llb[b];
    a.ll[p] = s.ll; a.lp[p] = s.lp; a.accept[p] = 0;
    if (r < R - 1) a.mc_accept[chain * (R - 1) + r] = 0;
    // row 0 of the burn-in outputs
    if (a.store_pt) { a.ring_ll[p] = s.ll; a.ring_lp[p] = s.lp; }
    else if (r == R - 1) { a.ring_ll[chain] = s.ll; a.ring_lp[chain] = s.lp; }
    if (a.save_hot_draws) {
      for (int i = 0; i < D; ++i) a.ring_theta[(long long)i * P + p] = s.theta[i];
    } else if (r == R - 1) {
      for (int i = 0; i < D; ++i) a.ring_theta[(long long)i * a.chains + chain] = s.theta[i];
    }
  }
}

// barrier of the coupling / abort phases: per warp when every chain lives inside one warp (R | 32)
// and the kernel has no CTA-wide data sweep, so that the warps of a CTA run independently
__device__ __forceinline__ void drj_sync(bool warp_local) {
  if (warp_local) __syncwarp();
  else __syncthreads();
}
__device__ __forceinline__ int drj_sync_or(bool warp_local, int pred) {
  if (warp_local) {
    const int any = __any_sync(0xffffffffu, pred);
    __syncwarp();
    return any;
  }
  return __syncthreads_or(pred);
}

// ------------------------------------------------------------------ sweep kernel body
// n_iter iterations of: update every rung (Particle::update), store, Metropolis coupling.
//
// STREAM = variant for data too large to stay in shared memory: the particle state is parked in its
// HBM slots for the duration of every data sweep and reloaded after it, so that nothing but the
// proposal (a dozen values) is live across the sweep and the sweep's inner loop gets the whole
// register file for its software pipeline (with the state live ptxas squeezes the loop into the ~48
// remaining registers and the LDS->DFMA distance collapses; profiles/README.md).  Costs ~0.3 KB of
// HBM traffic per thread per sweep of n >= 4097 data points.
//
// TEAM = one thread-block cluster per chain (few chains x long data, see drj_sweep_item_team): every
// thread is a replica of rung (tid mod R) of the cluster's chain and runs the whole update redundantly;
// `valid` = the thread holds a real particle, `active` = it is the replica that writes to HBM.
template <class M, bool STREAM, int MODE, bool SMALL = false>
__device__ __forceinline__ void drj_sweep_body(const DrjKernelArgs& a) {
  static_assert(!SMALL || (MODE == DRJ_MODE_SOLO && !STREAM), "SMALL: one thread per particle, data resident");
  DrjCtaInit<M>::run(a);
  constexpr bool TEAM = MODE == DRJ_MODE_TEAM, GRID = MODE == DRJ_MODE_GRID;
  static_assert(!(STREAM && MODE != DRJ_MODE_SOLO), "TEAM / GRID replicas keep their state in registers");
  constexpr int D = M::D, B = M::NBLOCK;
  constexpr int UD = DRJ_UNROLL_N(D), UB = DRJ_UNROLL_N(B);
  const int R = a.rungs;
  const int tid = threadIdx.x;
  const int lc = tid / R, r = tid - lc * R;
  // GRID: every CTA holds all chains (thread tid < P = particle tid); CTA 0's replicas write to HBM
  const long long chain = TEAM ? (long long)drj_cluster_id_x() : (GRID ? (long long)lc : (long long)blockIdx.x * a.chains_per_cta + lc);
  const bool valid = TEAM || (GRID ? (chain < a.chains) : ((lc < a.chains_per_cta) && (chain < a.chains)));
  const bool active = TEAM ? (lc == 0 && drj_cluster_ctarank() == 0u) : (valid && (!GRID || blockIdx.x == 0));
  const long long cc = valid ? chain : 0;
  const long long p = cc * R + r;
  const long long P = a.P;
  const int base = TEAM ? 0 : lc * R;  // first thread of this chain in the CTA (TEAM: slot 0 serves all replicas)

  __shared__ double s_tmin[D], s_tmax[D];
  __shared__ int s_trans[D], s_free[D];
  __shared__ double s_x[DRJ_EXCH * DRJ_NT];  // rung-swap exchange buffer; s_x[0..NT) doubles as loglike
  __shared__ double s_logu[DRJ_NT];
  __shared__ unsigned s_swap[DRJ_NT / 2];  // swap outcome of every pair, one bit each, (R + 30) / 32 words per chain
  __shared__ double s_beta[DRJ_NT];

  for (int i = tid; i < R; i += blockDim.x) s_beta[i] = a.beta_raised[i];
  for (int i = tid; i < D; i += blockDim.x) {
    s_tmin[i] = a.theta_min[i]; s_tmax[i] = a.theta_max[i];
    s_trans[i] = a.trans_type[i]; s_free[i] = a.free_slot[i];
  }
  int mc_count = 0;  // swaps accepted at this rung's pair in this launch

  DrjTiles<M> T;
  T.template init<MODE>(a);
  __syncthreads();

  // ---- load state into registers
  DrjState<M> s;
  drj_load_state<M>(a, s, p);
#pragma unroll UB
  for (int b = 0; b < B; ++b) s.llb_p[b] = s.llb[b];
  const double beta = a.beta_raised[r];
  const double target = a.target_acceptance;
  const unsigned gchain_lo = (unsigned)(unsigned long long)(a.chain_offset + cc);
  const unsigned gchain_hi = (unsigned)((unsigned long long)(a.chain_offset + cc) >> 32);
  const unsigned k0 = (unsigned)a.seed, k1 = (unsigned)(a.seed >> 32);
  const bool do_couple = a.coupling_on && R > 1;
  const bool wl = MODE == DRJ_MODE_SOLO && a.warp_local != 0;
  const int lane = tid & 31;
  const bool bt = B == 1 && a.blocks_trivial != 0;  // (no reads of the block lists: they were 5 % of the n = 100 sweep's instructions)

  // draws computed ahead (a.draws): element (it, free slot, particle) at ((it * d_free + slot) * P + p), so consecutive
  // updates of this thread are P apart; the loads run one update ahead of their use
  // (SMALL: a kernel of its own for these runs -- compiled into the kernel of the large runs, which sits at its
  // 128-register cap, the values held across the update made it spill: n = 100 sweep -2.5 %)
  const bool draws = SMALL && a.draws != 0;
  long long de = p;
  double z_nx = 0.0, lu_nx = 0.0;
  if (draws) { z_nx = a.draw_z[de]; lu_nx = a.draw_lu[de]; }

  for (int it = 0; it < a.n_iter; ++it) {
    const int t = a.iter0 + it;  // global update index, 1-based
    int err = 0;
    const double* rp = a.rng_mode ? a.replay + (cc * a.replay_T + (t - 1)) * a.replay_U : nullptr;
    double clu = 0.0;
    if (draws && do_couple && r < R - 1) clu = a.draw_clu[((long long)it * a.chains + cc) * (R - 1) + r];

    // ================= Particle::update =================
    constexpr int UP = (D <= 8) ? D : 1;  // the update body itself is only unrolled for small d
#pragma unroll UP
    for (int i = 0; i < D; ++i) {
      const int fs = s_free[i];
      if (fs < 0) continue;  // skip_param: no RNG consumed (Particle.h:91-93)
      double zq, lu3;  // the standard normal draw of the proposal, log u of the accept test
      if (draws) {
        zq = z_nx; lu3 = lu_nx;
        de += P;
        z_nx = a.draw_z[de]; lu_nx = a.draw_lu[de];
      } else {
        double u1, u2, u3;
        if (a.rng_mode) {
          const double* q = rp + 3 * ((long long)r * a.d_free + fs);
          u1 = q[0]; u2 = q[1]; u3 = q[2];
        } else {
          DrjU4 w = drj_philox4x32_10(gchain_lo, (unsigned)t, ((unsigned)r << 16) | (unsigned)i, gchain_hi << 8, k0, k1);
          u1 = drj_u32_to_unif(w.x); u2 = drj_u32_to_unif(w.y); u3 = drj_u32_to_unif(w.z);
        }
        zq = drj_norm_rand(u1, u2);
        lu3 = log(u3);
      }
      // propose_phi
      const double bw = exp(s.logbw[i]);
      const double phi_p = drj_rnorm_from(s.phi[i], bw, zq);
      // phi_prop_to_theta_prop + get_adjustment
      const double tmin = s_tmin[i], tmax = s_tmax[i];
      double th_p, lo_p = 0.0, hi_p = 0.0, adj = 0.0;
      switch (s_trans[i]) {
        case 0: th_p = phi_p; break;
        case 1: th_p = tmax - exp(phi_p); hi_p = log(tmax - th_p); adj = hi_p - s.ahi[i]; break;
        case 2: th_p = exp(phi_p) + tmin; lo_p = log(th_p - tmin); adj = lo_p - s.alo[i]; break;
        default: {
          const double e = exp(phi_p);
          th_p = __dadd_rn(__dmul_rn(tmax, e), tmin) / (1.0 + e);  // (no FMA: the reference rounds the product, Particle.cpp:68)
          hi_p = log(tmax - th_p); lo_p = log(th_p - tmin);
          adj = hi_p + lo_p - s.ahi[i] - s.alo[i];
        } break;
      }
      const double th_cur = s.theta[i], alo_cur = s.alo[i];
      s.theta[i] = th_p;  // s.theta is theta_prop (and s.alo its Jacobian logarithms) from here to the accept/reject
      s.alo[i] = lo_p;
      // likelihood of the blocks this parameter belongs to (Particle.h:107-112)
      const int jb = bt ? 0 : a.block_ptr[i], je = bt ? 1 : a.block_ptr[i + 1];
      for (int j = jb; j < je; ++j) {
        const int b = bt ? 1 : a.block_idx[j];
        if constexpr (STREAM) {
          typename M::Consts k;
          const bool fast = DrjHints<M>::prepare(s.theta, a.misc, b, k, s.alo, a.lo_is_log);
          const int item = M::datum_item(b);
          const long long n = a.data.len(item);
          if (active) drj_store_state<M>(a, s, p);                          // park (theta slot i holds theta')
          const double acc = drj_sweep_item_inl<M>(T, k, a.data.item(item), n);  // uniform: all threads sweep
          drj_load_state<M>(a, s, p);
          s.llb_p[B == 1 ? 0 : b - 1] = fast ? M::finish(k, acc, n) : M::loglike(s.theta, a.data, a.misc, b);
        } else {
          s.llb_p[B == 1 ? 0 : b - 1] = drj_eval_loglike<M, MODE>(T, a, s.theta, b, (int)p, s.alo);
        }
      }
      double ll_p = 0.0;
#pragma unroll UB
      for (int b = 0; b < B; ++b) ll_p += s.llb_p[b];  // sequential, index order (misc.h:20-27)
      const double lp_p = DrjHints<M>::logprior(s.theta, a.misc, s.alo, a.lo_is_log);
      // NaN / +Inf guard (Particle.h:121-128)
      if (err == 0 && valid) {
        if (isnan(ll_p) || ll_p == DRJ_INF) err = DRJ_ERR_LOGLIKE_BAD;
        else if (isnan(lp_p) || lp_p == DRJ_INF) err = DRJ_ERR_LOGPRIOR_BAD;
        if (err && active) drj_report(a, err, it, chain, r, i, p, s.theta, D);
      }
      // Metropolis-Hastings (Particle.h:131-139)
      // (explicitly rounded product and sums, left to right as the reference's expression: an FMA contraction here
      // could flip a decision that sits within an ulp of log u)
      double MH;
      if (beta == 0.0) MH = __dadd_rn(lp_p - s.lp, adj);
      else MH = __dadd_rn(__dadd_rn(__dmul_rn(beta, ll_p - s.ll), lp_p - s.lp), adj);
      const bool accept = (lu3 < MH) && (err == 0);
      const double gain = rsqrt((double)s.bwi[i]);  // bw_stepsize = 1
      // accept / reject as selects, not as a branch: the lanes of a warp decide differently at every update, and a
      // divergent branch here made every warp run both sides (n = 100 sweep: 26.7 of 32 lanes active, round 1)
#ifdef DRJ_ACCEPT_BRANCH
      if (accept) {
        s.phi[i] = phi_p; s.ahi[i] = hi_p;
        for (int j = jb; j < je; ++j) {
          const int b = (B == 1) ? 0 : a.block_idx[j] - 1;
          s.llb[b] = s.llb_p[b];
        }
        s.ll = ll_p; s.lp = lp_p;
        s.logbw[i] = fma(1.0 - target, gain, s.logbw[i]);
        s.acc++;
      } else {
        s.theta[i] = th_cur; s.alo[i] = alo_cur;
        for (int j = jb; j < je; ++j) {
          const int b = (B == 1) ? 0 : a.block_idx[j] - 1;
          s.llb_p[b] = s.llb[b];
        }
        s.logbw[i] = fma(-target, gain, s.logbw[i]);
      }
      s.bwi[i]++;
#else
      s.phi[i] = accept ? phi_p : s.phi[i];
      s.alo[i] = accept ? lo_p : alo_cur;
      s.ahi[i] = accept ? hi_p : s.ahi[i];
      s.theta[i] = accept ? th_p : th_cur;
      for (int j = jb; j < je; ++j) {
        const int b = (B == 1) ? 0 : a.block_idx[j] - 1;
        const double v = accept ? s.llb_p[b] : s.llb[b];
        s.llb[b] = v;
        s.llb_p[b] = v;
      }
      s.ll = accept ? ll_p : s.ll;
      s.lp = accept ? lp_p : s.lp;
      // Robbins-Monro on the log scale: + (1 - a*) / sqrt(index) on acceptance (Particle.h:153), - a* / sqrt(index) otherwise (:165)
      s.logbw[i] = fma(accept ? (1.0 - target) : -target, gain, s.logbw[i]);
      s.acc += accept ? 1 : 0;
      s.bwi[i]++;
#endif
    }

    // ================= store (before coupling, main.cpp:165-173) =================
    if (active) {
      if (a.store_pt) {
        a.ring_ll[(long long)it * P + p] = s.ll;
        a.ring_lp[(long long)it * P + p] = s.lp;
      } else if (r == R - 1) {
        a.ring_ll[(long long)it * a.chains + chain] = s.ll;
        a.ring_lp[(long long)it * a.chains + chain] = s.lp;
      }
      if (a.save_hot_draws) {
#pragma unroll UD
        for (int i = 0; i < D; ++i) a.ring_theta[((long long)it * D + i) * P + p] = s.theta[i];
      } else if (r == R - 1) {
#pragma unroll UD
        for (int i = 0; i < D; ++i) a.ring_theta[((long long)it * D + i) * a.chains + chain] = s.theta[i];
      }
    }

    // ================= Metropolis coupling (main.cpp:282-344) =================
    int src = r;
    bool couple_serial = do_couple;
    if (SMALL && do_couple && a.couple_par) {
      // Runs too small to fill the machine (latency-bound), R <= 32.  The serial walk below costs ~26 dependent
      // instructions per pair in ONE lane (K2, 20 rungs: 21 % of the run).  The replica carried into pair i came from
      // some rung j <= i: thread i takes the pair's decision for EVERY such j at once (independent instructions, all
      // lanes busy), one bit each -- the same rounded expression as the walk, main.cpp:300-308 -- and the walk itself
      // is three integer instructions per pair on those masks, done by every thread for its own chain (no walker lane,
      // no swap mask to publish).
      couple_serial = false;
      unsigned* const s_dec = reinterpret_cast<unsigned*>(s_logu);
      s_x[tid] = s.ll;
      drj_sync(wl);
      if (r < R - 1) {
        double lu;
        if (draws) lu = clu;
        else {
          double u;
          if (a.rng_mode) u = rp[3LL * a.d_free * R + r];
          else {
            DrjU4 w = drj_philox4x32_10(gchain_lo, (unsigned)t, (unsigned)r, 1u | (gchain_hi << 8), k0, k1);
            u = drj_u32_to_unif(w.x);
          }
          lu = log(u);
        }
        const double ll2 = s_x[tid + 1], b1 = beta, b2 = s_beta[r + 1];
        const double q1 = __dmul_rn(ll2, b1), q2 = __dmul_rn(ll2, b2);
        unsigned m = 0u;
#pragma unroll 4
        for (int j = 0; j <= r; ++j) {
          const double cur_ll = s_x[base + j];
          const double t1 = __dmul_rn(cur_ll, b2), t2 = __dmul_rn(cur_ll, b1);
          const double acceptance = (b1 == 0.0) ? __dsub_rn(t1, q2) : __dsub_rn(__dadd_rn(q1, t1), __dadd_rn(t2, q2));
          m |= (lu < acceptance ? 1u : 0u) << j;
        }
        s_dec[tid] = m;
      }
      drj_sync(wl);
      unsigned sw = 0u;
      {
        int j = 0;
#pragma unroll 4
        for (int i = 0; i < R - 1; ++i) {
          const unsigned bit = (s_dec[base + i] >> j) & 1u;
          sw |= bit << i;
          j = bit ? j : i + 1;
        }
      }
      if (valid) {
        if (r < R - 1 && ((sw >> r) & 1u)) {
          src = r + 1;
          ++mc_count;
        } else {
          const unsigned zeros = ~sw & ((1u << r) - 1u);  // pairs below r that did not swap
          src = zeros ? 32 - __clz(zeros) : 0;
        }
      }
    }
    if (couple_serial) {
      // parallel part: every rung publishes its loglike, the two products of the pair BELOW it that do not depend on
      // the outcome of earlier swaps (loglike2 * beta1, loglike2 * beta2: rung i+1 is untouched until pair i is
      // decided), and the pair's log u
      if (!TEAM || tid < R) {
        s_x[tid] = s.ll;
        if (r > 0) {
          s_x[DRJ_NT + tid] = __dmul_rn(s.ll, s_beta[r - 1]);
          s_x[2 * DRJ_NT + tid] = __dmul_rn(s.ll, beta);
        }
      }
      if (r < R - 1 && (!TEAM || tid < R)) {
        if (draws) s_logu[tid] = clu;
        else {
          double u;
          if (a.rng_mode) u = rp[3LL * a.d_free * R + r];
          else {
            DrjU4 w = drj_philox4x32_10(gchain_lo, (unsigned)t, (unsigned)r, 1u | (gchain_hi << 8), k0, k1);
            u = drj_u32_to_unif(w.x);
          }
          s_logu[tid] = log(u);
        }
      }
      drj_sync(wl);
      // sequential part hot -> cold (the swap of pair i sees the outcome of pair i-1: the replica that moves up carries
      // its loglike), one chain per LANE of the first warp(s).  The walker only DECIDES: per pair two products with the
      // carried loglike, three rounded sums, a compare; the outcome is one bit of a mask.  (Round 1 wrote the source
      // rung and the swap count of every pair from inside this loop: ~70 instructions per pair at 4/32 lanes, 17 % of
      // all instructions issued by the n = 100 sweep with 32 rungs, and the other warps of the CTA waiting for it.)
      // warp_local: the first 32/R lanes of each warp walk that warp's chains.
      const int W = (R + 30) >> 5;  // mask words per chain (R - 1 pairs)
      const int wb = wl ? ((tid & ~31) + lane * R) : tid * R;
      const int wlc = wb / R;
      const bool walker = TEAM ? (tid == 0)
                               : (GRID ? (tid < a.chains)
                                       : ((wl ? (lane < 32 / R) : true) && wlc < a.chains_per_cta &&
                                          (long long)blockIdx.x * a.chains_per_cta + wlc < a.chains));
      if (walker) {
        unsigned m = 0u;
        double cur_ll = s_x[wb];
        double b1 = s_beta[0];
#pragma unroll 4
        for (int i = 0; i < R - 1; ++i) {
          const double ll2 = s_x[wb + i + 1], q1 = s_x[DRJ_NT + wb + i + 1], q2 = s_x[2 * DRJ_NT + wb + i + 1];
          const double b2 = s_beta[i + 1], lu = s_logu[wb + i];
          // main.cpp:300-308, every product and sum rounded (no FMA contraction in a decision):
          //   (loglike2 * beta1 + loglike1 * beta2) - (loglike1 * beta1 + loglike2 * beta2)
          const double t1 = __dmul_rn(cur_ll, b2), t2 = __dmul_rn(cur_ll, b1);
          const double acceptance = (b1 == 0.0) ? __dsub_rn(t1, q2) : __dsub_rn(__dadd_rn(q1, t1), __dadd_rn(t2, q2));
          const bool swap = lu < acceptance;
          m |= (swap ? 1u : 0u) << (i & 31);
          cur_ll = swap ? cur_ll : ll2;  // rejected: the replica of rung i+1 becomes the one that may move up
          b1 = b2;
          if ((i & 31) == 31) { s_swap[wlc * W + (i >> 5)] = m; m = 0u; }
        }
        if ((R - 1) & 31) s_swap[wlc * W + ((R - 2) >> 5)] = m;
      }
      drj_sync(wl);
      // parallel part: rung r receives the replica of rung r+1 if pair r swapped, else the replica that was being
      // carried up: the one from below the run of swapped pairs that ends at pair r-1
      if (valid) {
        const unsigned* sw = s_swap + (TEAM ? 0 : lc) * W;
        if (r < R - 1 && ((sw[r >> 5] >> (r & 31)) & 1u)) {
          src = r + 1;
          ++mc_count;
        } else {
          // highest pair below r that did NOT swap, + 1 (0 if every pair below swapped): one __clz per mask word, not a
          // walk down the run -- on a ladder whose hot rungs swap almost always the runs are as long as the ladder
          src = 0;
          for (int pos = r; pos > 0;) {
            const int w = (pos - 1) >> 5, nb = pos - (w << 5);  // bits [0, nb) of word w lie below `pos`
            const unsigned zeros = ~sw[w] & (nb == 32 ? 0xffffffffu : ((1u << nb) - 1u));
            if (zeros) { src = (w << 5) + 32 - __clz(zeros); break; }
            pos = w << 5;
          }
        }
      }
    }
    if (do_couple) {
      const int from = base + src;
      if (drj_sync_or(wl, src != r)) {
        // replica-owned state moves: theta, phi, Jacobian logs, loglike_block, loglike, logprior.
        // rung-owned state stays: beta, bw, bw_index, accept_count (main.cpp:317-336).
        constexpr int NV = 4 * D + B + 2;
        constexpr int UX = (D <= 8 && B <= 8) ? 16 : 1;
#pragma unroll UX
        for (int v0 = 0; v0 < NV; v0 += DRJ_EXCH) {
#pragma unroll
          for (int e = 0; e < DRJ_EXCH; ++e) {
            const int v = v0 + e;
            if (v < NV) {
              double val;
              if (v < D) val = s.theta[v];
              else if (v < 2 * D) val = s.phi[v - D];
              else if (v < 3 * D) val = s.alo[v - 2 * D];
              else if (v < 4 * D) val = s.ahi[v - 3 * D];
              else if (v < 4 * D + B) val = s.llb[v - 4 * D];
              else if (v == 4 * D + B) val = s.ll;
              else val = s.lp;
              if (!TEAM || tid < R) s_x[e * DRJ_NT + tid] = val;
            }
          }
          drj_sync(wl);
#pragma unroll
          for (int e = 0; e < DRJ_EXCH; ++e) {
            const int v = v0 + e;
            if (v < NV) {
              const double val = s_x[e * DRJ_NT + from];
              if (v < D) s.theta[v] = val;
              else if (v < 2 * D) s.phi[v - D] = val;
              else if (v < 3 * D) s.alo[v - 2 * D] = val;
              else if (v < 4 * D) s.ahi[v - 3 * D] = val;
              else if (v < 4 * D + B) { s.llb[v - 4 * D] = val; s.llb_p[v - 4 * D] = val; }
              else if (v == 4 * D + B) s.ll = val;
              else s.lp = val;
            }
          }
          drj_sync(wl);
        }
      }
    }

    // ================= abort on error (Rcpp::stop in the reference) =================
    // one barrier: own error, or (seen by thread 0) an error reported by any other CTA
    // (TEAM / GRID: the decision must be the same in every CTA of the cluster / grid, which meet at a barrier
    // in the next sweep -- own errors only, they are identical in all replicas)
    int stop = err;
    if (MODE == DRJ_MODE_SOLO && (wl ? lane == 0 : tid == 0) && *((volatile unsigned long long*)a.err_key) != ~0ULL) stop = 1;
    if (drj_sync_or(wl, stop)) break;
  }
  drj_sync(wl);

  // ---- write state back
  if (active) {
    drj_store_state<M>(a, s, p);
    if (do_couple && r < R - 1) a.mc_accept[chain * (R - 1) + r] += mc_count;
  }
}

// ------------------------------------------------------------------ kernels
template <class M>
__global__ void __launch_bounds__(DRJ_NT) drj_init_kernel(const __grid_constant__ DrjKernelArgs a) {
  drj_init_body<M, DRJ_MODE_SOLO>(a);
}
template <class M>
__global__ void __launch_bounds__(DRJ_NT) drj_init_chain_kernel(const __grid_constant__ DrjKernelArgs a) {
  drj_init_chain_body<M>(a);
}
template <class M, bool STREAM>
__global__ void __launch_bounds__(DRJ_NT, (STREAM ? 2 : DRJ_MINB)) drj_sweep_kernel(const __grid_constant__ DrjKernelArgs a) {
  drj_sweep_body<M, STREAM, DRJ_MODE_SOLO>(a);
}
// runs too small to fill the machine (at most two warps per scheduler): draws computed ahead, coupling decisions taken
// for every possible carried replica at once (drj_sweep_body, SMALL); one CTA per SM is plenty, so no register cap
template <class M>
__global__ void __launch_bounds__(DRJ_NT, 1) drj_sweep_kernel_small(const __grid_constant__ DrjKernelArgs a) {
  drj_sweep_body<M, false, DRJ_MODE_SOLO, true>(a);
}
// cluster-per-chain variants (launched with a cluster dimension of a.team_ctas; datum-form models only)
template <class M>
__global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_init_kernel_team(const __grid_constant__ DrjKernelArgs a) {
  drj_init_body<M, DRJ_MODE_TEAM>(a);
}
template <class M>
__global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_sweep_kernel_team(const __grid_constant__ DrjKernelArgs a) {
  drj_sweep_body<M, false, DRJ_MODE_TEAM>(a);
}
// grid-per-run variants (cooperative launch: all CTAs co-resident; datum-form models only)
template <class M>
__global__ void __launch_bounds__(DRJ_NT, 1) drj_init_kernel_grid(const __grid_constant__ DrjKernelArgs a) {
  drj_init_body<M, DRJ_MODE_GRID>(a);
}
template <class M>
__global__ void __launch_bounds__(DRJ_NT, 1) drj_sweep_kernel_grid(const __grid_constant__ DrjKernelArgs a) {
  drj_sweep_body<M, false, DRJ_MODE_GRID>(a);
}
// addresses of the variants that only exist for datum-form models
template <class M, bool DF = M::DATUM_FORM>
struct DrjStreamKernel {
  static const void* get() { return nullptr; }
  static const void* team_init() { return nullptr; }
  static const void* team_sweep() { return nullptr; }
  static const void* grid_init() { return nullptr; }
  static const void* grid_sweep() { return nullptr; }
};
template <class M>
struct DrjStreamKernel<M, true> {
  static const void* get() { return (const void*)drj_sweep_kernel<M, true>; }
  static const void* team_init() { return (const void*)drj_init_kernel_team<M>; }
  static const void* team_sweep() { return (const void*)drj_sweep_kernel_team<M>; }
  static const void* grid_init() { return (const void*)drj_init_kernel_grid<M>; }
  static const void* grid_sweep() { return (const void*)drj_sweep_kernel_grid<M>; }
};
