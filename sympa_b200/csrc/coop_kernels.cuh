// Warp-cooperative kernel for the larger matrix sizes (upper half space, n >= 5): a pair is worked
// on by G = ceil(n/2) lanes, 32/G pairs per warp, every matrix in shared memory (coop_math.cuh).
// One warp per CTA: the only barrier needed is __syncwarp().
#pragma once
#include "coop_math.cuh"
#include "pair_kernels.cuh"

namespace sympa {

#ifdef SYMPA_PAIR_KERNELS_IMPL

struct WarpExec {
  int g;
  bool active;
  __device__ __forceinline__ int first() const { return active ? g : 0; }
  __device__ __forceinline__ int last() const { return active ? g + 1 : 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(0xffffffffu, p && active) != 0; }

  // Register-resident one-sided Jacobi: each lane keeps two columns in registers, rotates them, and between two
  // rounds passes ONE of them to its right neighbour in the group's ring (one-directional ring ordering).  Which one -
  // the column it has just received (bit 1) or the one it kept (bit 0) - is a fixed schedule per (round, lane) found
  // by exhaustive search (ring_send_masks): N - 1 rounds in which every pair of columns meets exactly once, from
  // any starting arrangement, so sweep after sweep runs the same code.  Every lane of the warp executes the shuffles
  // (inactive lanes carry garbage that nobody reads).
  // Round 1 and most of round 2 used the Brent-Luk ordering here: two columns leave every lane per round (top row one
  // lane up, bottom row one lane down) and the first / last lane of a group need selects on both: 4 SHFL + 6 FSEL
  // per exchanged double against 2 SHFL + 4 SEL now - the exchange was 26 % of the instructions of the n = 10 kernel.
  // Measured: n = 10 11.19 -> 10.95 ms per 2^19 pairs, n = 9 8.79 -> 8.72, n = 8 5.14 -> 5.12: two per cent - the rounds
  // are bound by the latency of the dependent chain dots -> rotation parameters -> update at 1.25 warps per scheduler,
  // not by the instructions of the exchange.
  // (Round 2 also tried the Brent-Luk exchange through shared memory - STS.128 / LDS.128 into per-lane slots of the
  // pair's idle G buffers, ~110 instructions per round - and measured it SLOWER: n = 10 13.6 ms against 11.2 ms per 2^19
  // pairs: two __syncwarp() and a store -> load round trip per round cost more at 1.25 warps per scheduler than
  // shuffles, which pipeline.)
  template <int N, bool IS_REAL>
  __device__ __forceinline__ int jacobi(double* gr, double* gi, double*) const {
    typedef coop::LayoutT<N, 2> L;
    constexpr int NP = L::NP, LD = L::LD, G = L::G;
    constexpr unsigned kFull = 0xffffffffu;
    constexpr unsigned long long kSched = coop::ring_send_masks<G>();
    const int src = (threadIdx.x & 31) - g + (g == 0 ? G - 1 : g - 1);   // left neighbour in the group's ring
    double tr[N], ti[N], br[N], bi[N];
    int tid = 2 * g, bid = 2 * g + 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      tr[i] = gr[tid * LD + i];
      ti[i] = IS_REAL ? 0.0 : gi[tid * LD + i];
      br[i] = bid < N ? gr[bid * LD + i] : 0.0;
      bi[i] = (bid < N && !IS_REAL) ? gi[bid * LD + i] : 0.0;
    }
    __syncwarp();
    int sweep = 0;
    bool done = !active;  // this pair's group has converged: its columns are frozen from here on
#pragma unroll 1
    for (; sweep < kMaxSweeps; ++sweep) {
      unsigned conv = 0u;
      double tn = coop::column_norm2<N, IS_REAL>(tr, ti);  // squared norms, carried through the sweep
      double bn = coop::column_norm2<N, IS_REAL>(br, bi);
// (the round loop stays rolled: unrolled by 3 the n = 10 kernel took 13.97 ms per 2^19 pairs instead of 10.96, fully
// unrolled 16.25 ms - five warps per SM in different phases of 134 kB of code live off the instruction cache)
#ifndef SY_COOP_ROUND_UNROLL
#define SY_COOP_ROUND_UNROLL 1
#endif
      constexpr int kRoundUnroll = SY_COOP_ROUND_UNROLL;
#pragma unroll kRoundUnroll
      for (int r = 0; r < NP - 1; ++r) {
        if (!done) conv |= coop::rotate_columns<N, IS_REAL>(tr, ti, br, bi, &tn, &bn);
        if (G > 1) {
          // bit r * G + g of the schedule: 1 = pass on the column received last (bottom), 0 = the kept one (top)
          const bool pass_b = ((kSched >> (r * G + g)) & 1ull) != 0ull;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            const double send_r = pass_b ? br[i] : tr[i];
            tr[i] = pass_b ? tr[i] : br[i];
            br[i] = __shfl_sync(kFull, send_r, src);
            if (!IS_REAL) {
              const double send_i = pass_b ? bi[i] : ti[i];
              ti[i] = pass_b ? ti[i] : bi[i];
              bi[i] = __shfl_sync(kFull, send_i, src);
            }
          }
          {
            const double send_n = pass_b ? bn : tn;
            tn = pass_b ? tn : bn;
            bn = __shfl_sync(kFull, send_n, src);
          }
          const int send_id = pass_b ? bid : tid;
          tid = pass_b ? tid : bid;
          bid = __shfl_sync(kFull, send_id, src);
        }
      }
      // the flags are combined over the lanes of THIS pair's group only: a pair's sweep count, and
      // therefore its result, does not depend on which other pairs share the warp.  The loop runs
      // until every group of the warp is done; converged groups keep shuffling (the ring must stay
      // in step) but no longer rotate.
      unsigned gconv = 0u;
#pragma unroll
      for (int k = 0; k < G; ++k) gconv |= __shfl_sync(kFull, conv, (threadIdx.x & 31) - g + k);
      done = done || loc::jacobi_sweep_is_last(gconv);
      if (__all_sync(kFull, done)) {
        ++sweep;
        break;
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (tid < N) gr[tid * LD + i] = tr[i];
        if (bid < N) gr[bid * LD + i] = br[i];
        if (!IS_REAL) {
          if (tid < N) gi[tid * LD + i] = ti[i];
          if (bid < N) gi[bid * LD + i] = bi[i];
        }
      }
    }
    __syncwarp();
    return sweep;
  }
};

template <int N>
struct CoopCfg {
  typedef coop::Layout<N> L;
  static constexpr int G = L::G;
  static constexpr int PW = 32 / G;  // pairs per warp
  static constexpr int kSmemBytes = PW * L::kDoubles * (int)sizeof(double);
};

// per-kind view of a pair's shared-memory region: size, where the scalars are, where the four
// gradient blocks (re/im of operand 1 and 2) are and with which sign
template <int N, int KIND>
struct CoopTraits;
template <int N>
struct CoopTraits<N, kUpper> {
  typedef coop::Layout<N> L;
  static constexpr int kDoubles = L::kDoubles, DIST = L::DIST, VS = L::VS, FLAG = L::FLAG;
  static constexpr int R1 = L::GX2, I1 = L::GY1, R2 = L::GX2, I2 = L::GY2;
  static constexpr double kSignR1 = -1.0;  // d/dX1 = -d/dX2
};
template <int N>
struct CoopTraits<N, kSpd> {
  typedef coop::SpdLayout<N> L;
  static constexpr int kDoubles = L::kDoubles, DIST = L::DIST, VS = L::VS, FLAG = L::FLAG;
  static constexpr int R1 = L::SPD_GX, I1 = L::SPD_GX, R2 = L::SPD_GY, I2 = L::SPD_GY;
  static constexpr double kSignR1 = 1.0;
};
template <int N>
struct CoopTraits<N, kBounded> {
  typedef coop::BoundedLayout<N> L;
  static constexpr int kDoubles = L::kDoubles, DIST = L::DIST, VS = L::VS, FLAG = L::FLAG;
  static constexpr int R1 = L::G1R, I1 = L::G1I, R2 = L::G2R, I2 = L::G2I;
  static constexpr double kSignR1 = 1.0;
};

// writes the unit gradients (MODE fwd+save) or the scaled scatter-add (MODE step) from the results the
// pair mathematics left in shared memory (symmetrised here)
template <int N, int KIND, int MODE>
__device__ __forceinline__ void emit_gradients(const PairArgs& a, const double* sm, int g, int64_t p, int64_t i1,
                                               int64_t i2, double dist, double* loss_acc, double* gscale_acc) {
  typedef CoopTraits<N, KIND> TR;
  typedef coop::Layout<N> L;
  constexpr int G = L::G;
  constexpr int NN = N * N;
  constexpr int PER = (KIND == kSpd ? 1 : 2) * NN;
  constexpr int LD = L::LD;
  const double* r1 = sm + TR::R1;
  const double* r2 = sm + TR::R2;
  const double s1 = TR::kSignR1;
  const double* m1 = sm + TR::I1;
  const double* m2 = sm + TR::I2;
  double scale = 1.0;
  double* o1;
  double* o2;
  constexpr int TRI = N * (N + 1) / 2;                    // packed lower triangle of one block
  constexpr int PS = (KIND == kSpd ? 1 : 2) * TRI;        // saved state of one point (state_doubles)
  if (MODE == kModeFwdSave) {
    o1 = a.gz1 + p * PS;
    o2 = a.gz2 + p * PS;
  } else {  // fused distortion step: L_p = |(s d / g)^2 - 1|  (losses.py:16-19, model.py:30)
    const double gd = __ldg(a.graph_dist + p);
    const double r = a.scale * dist / gd;
    const double er = r * r - 1.0;
    const double sg = er > 0.0 ? 1.0 : (er < 0.0 ? -1.0 : 0.0);
    const double dl_dr = sg * 2.0 * r;
    scale = dl_dr * a.scale / gd;
    if (g == 0) {
      *loss_acc += fabs(er);
      *gscale_acc += dl_dr * dist / gd;
    }
    o1 = a.grad_table + i1 * PER;
    o2 = a.grad_table + i2 * PER;
  }
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N, s = i * LD + j, t = j * LD + i;
    if (MODE == kModeFwdSave && j > i) continue;          // the saved state keeps the lower triangle only
    const int pk = i * (i + 1) / 2 + j;
    const double v1 = scale * s1 * 0.5 * (r1[s] + r1[t]);
    const double v2 = scale * 0.5 * (r2[s] + r2[t]);
    if (MODE == kModeFwdSave) {
      o1[pk] = v1;
      o2[pk] = v2;
    } else {
      atomicAdd(o1 + e, v1);
      atomicAdd(o2 + e, v2);
    }
    if (KIND != kSpd) {
      const double w1 = scale * 0.5 * (m1[s] + m1[t]);
      const double w2 = scale * 0.5 * (m2[s] + m2[t]);
      if (MODE == kModeFwdSave) {
        o1[TRI + pk] = w1;
        o2[TRI + pk] = w2;
      } else {
        atomicAdd(o1 + NN + e, w1);
        atomicAdd(o2 + NN + e, w2);
      }
    }
  }
}

template <int N, int KIND, int MODE>
__global__ void __launch_bounds__(32) coop_kernel(const PairArgs a) {
  typedef CoopTraits<N, KIND> L;
  constexpr int G = CoopCfg<N>::G;
  constexpr int PW = CoopCfg<N>::PW;
  constexpr int NN = N * N;
  constexpr int PER = (KIND == kSpd ? 1 : 2) * NN;
  constexpr bool GRAD = MODE != kModeFwd;
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int slot = lane / G;
  const int g = lane - slot * G;
  double* sm = smem + (slot < PW ? slot : 0) * L::kDoubles;
  unsigned st = 0;
  double loss_acc = 0.0, gscale_acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * PW;
  for (int64_t base = (int64_t)blockIdx.x * PW; base < a.num_pairs; base += stride) {
    const int64_t p = base + slot;
    bool active = slot < PW && p < a.num_pairs;
    const double* p1 = nullptr;
    const double* p2 = nullptr;
    int64_t i1 = 0, i2 = 0;
    if (active) {
      if (a.idx != nullptr) {
        i1 = __ldg(a.idx + 2 * p);
        i2 = __ldg(a.idx + 2 * p + 1);
        if (i1 < 0 || i1 >= a.num_rows || i2 < 0 || i2 >= a.num_rows) {
          st |= kStatusBadIndex;
          if (g == 0 && a.dist_out) a.dist_out[p] = 0.0;
          active = false;
        } else {
          p1 = a.table + i1 * PER;
          p2 = a.table + i2 * PER;
        }
      } else {
        p1 = a.z1 + p * PER;
        p2 = a.z2 + p * PER;
      }
    }
    // L2 prefetch of the rows of this lane's NEXT pair (the table gather is a dependent chain
    // idx -> row address -> data whose latency is exposed at one warp per scheduler)
    {
      const int64_t pn = p + stride;
      if (slot < PW && pn < a.num_pairs) {
        const double* q1;
        const double* q2;
        bool okn = true;
        if (a.idx != nullptr) {
          const int64_t j1 = __ldg(a.idx + 2 * pn), j2 = __ldg(a.idx + 2 * pn + 1);
          okn = j1 >= 0 && j1 < a.num_rows && j2 >= 0 && j2 < a.num_rows;
          q1 = a.table + (okn ? j1 : 0) * PER;
          q2 = a.table + (okn ? j2 : 0) * PER;
        } else {
          q1 = a.z1 + pn * PER;
          q2 = a.z2 + pn * PER;
        }
        if (okn) {
          for (int off = g * 16; off < PER; off += G * 16) {  // one 128-byte line per step
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + off));
          }
        }
      }
    }
    WarpExec ex{g, active};
    if (KIND == kSpd)
      coop::spd_pair<N, GRAD>(ex, sm, p1, p2);
    else if (KIND == kBounded)
      coop::bounded_pair<N, GRAD>(ex, sm, p1, p2, a.metric, a.wsum_w);
    else
      coop::upper_pair<N, GRAD>(ex, sm, p1, p2, a.metric, a.wsum_w);
    if (active) {
      const double dist = sm[L::DIST];
      if (g == 0) {
        st |= (unsigned)sm[L::FLAG];
        if (a.dist_out) a.dist_out[p] = dist;
      }
      if (a.vvd_out) {
        for (int k = g; k < N; k += G) a.vvd_out[p * N + k] = sm[L::VS + k];
      }
      if (GRAD) {
        if (MODE == kModeStep && KIND != kSpd && a.grad_wsum_w != nullptr && a.metric == kWsum) {
          const double gd = __ldg(a.graph_dist + p);
          const double r = a.scale * dist / gd;
          const double er = r * r - 1.0;
          const double dl_dd = (er > 0.0 ? 1.0 : (er < 0.0 ? -1.0 : 0.0)) * 2.0 * r * a.scale / gd;
          for (int k = g; k < N; k += G)
            if (a.wsum_w[k] > 0.0) atomicAdd(a.grad_wsum_w + k, dl_dd * sm[L::VS + k]);
        }
        emit_gradients<N, KIND, MODE>(a, sm, g, p, i1, i2, dist, &loss_acc, &gscale_acc);
      }
    }
    __syncwarp();
  }
  if (MODE == kModeStep) {
    loss_acc = warp_sum(loss_acc);
    gscale_acc = warp_sum(gscale_acc);
    if (lane == 0) {
      if (a.loss_out) atomicAdd(a.loss_out, loss_acc);
      if (a.grad_scale) atomicAdd(a.grad_scale, gscale_acc);
    }
  }
  st = __reduce_or_sync(0xffffffffu, st);
  if (st != 0 && lane == 0 && a.status != nullptr) atomicOr(a.status, st);
}

// ------------------------------------------------------------------------------------------------
// Split path: the same pipeline as three kernels with the per-pair state parked in global scratch
// (L2-resident for the chunk sizes used).  The Jacobi kernel then only needs the two halves of G in
// shared memory, so four to five times more pairs are resident per SM while the latency-bound
// rotation chains run - which is where ~70 % of the FP64 work is.
struct ChunkArgs {
  double* scratch;
  int64_t cap;    // scratch capacity in pairs
  int64_t base;   // first pair of this chunk
  int64_t count;  // pairs in this chunk (<= cap)
};

template <int N>
__device__ __forceinline__ bool locate_pair(const PairArgs& a, int64_t p, const double** p1, const double** p2,
                                            int64_t* i1, int64_t* i2) {
  constexpr int PER = 2 * N * N;
  if (a.idx != nullptr) {
    *i1 = __ldg(a.idx + 2 * p);
    *i2 = __ldg(a.idx + 2 * p + 1);
    if (*i1 < 0 || *i1 >= a.num_rows || *i2 < 0 || *i2 >= a.num_rows) return false;
    *p1 = a.table + *i1 * PER;
    *p2 = a.table + *i2 * PER;
  } else {
    *p1 = a.z1 + p * PER;
    *p2 = a.z2 + p * PER;
  }
  return true;
}

template <int N>
__global__ void __launch_bounds__(32) coop_prologue_kernel(const PairArgs a, const ChunkArgs c) {
  typedef coop::Layout<N> L;
  constexpr int G = CoopCfg<N>::G;
  constexpr int PW = CoopCfg<N>::PW;
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int slot = lane / G;
  const int g = lane - slot * G;
  double* sm = smem + (slot < PW ? slot : 0) * L::kDoubles;
  unsigned st = 0;
  const int64_t stride = (int64_t)gridDim.x * PW;
  for (int64_t base = (int64_t)blockIdx.x * PW; base < c.count; base += stride) {
    const int64_t q = base + slot;
    bool active = slot < PW && q < c.count;
    const double* p1 = nullptr;
    const double* p2 = nullptr;
    int64_t i1, i2;
    if (active && !locate_pair<N>(a, c.base + q, &p1, &p2, &i1, &i2)) {
      st |= kStatusBadIndex;
      if (g == 0 && a.dist_out) a.dist_out[c.base + q] = 0.0;
      active = false;
    }
    WarpExec ex{g, active};
    coop::split_prologue<N>(ex, sm, p1, p2, c.scratch, c.cap, q);
    if (active && g == 0 && sm[L::FLAG] != 0.0) st |= kStatusNotPD;
    __syncwarp();
  }
  st = __reduce_or_sync(0xffffffffu, st);
  if (st != 0 && lane == 0 && a.status != nullptr) atomicOr(a.status, st);
}

template <int N, bool GRAD>
__global__ void __launch_bounds__(32) coop_spectrum_kernel(const PairArgs a, const ChunkArgs c) {
  typedef coop::LayoutJ<N> L;
  constexpr int G = CoopCfg<N>::G;
  constexpr int PW = CoopCfg<N>::PW;
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int slot = lane / G;
  const int g = lane - slot * G;
  double* sm = smem + (slot < PW ? slot : 0) * L::kDoubles;
  unsigned st = 0;
  const int64_t stride = (int64_t)gridDim.x * PW;
  for (int64_t base = (int64_t)blockIdx.x * PW; base < c.count; base += stride) {
    const int64_t q = base + slot;
    const int64_t p = c.base + q;
    bool active = slot < PW && q < c.count;
    if (active && a.idx != nullptr) {
      const int64_t i1 = __ldg(a.idx + 2 * p), i2 = __ldg(a.idx + 2 * p + 1);
      if (i1 < 0 || i1 >= a.num_rows || i2 < 0 || i2 >= a.num_rows) active = false;
    }
    WarpExec ex{g, active};
    coop::split_spectrum<N, GRAD>(ex, sm, a.metric, a.wsum_w, c.scratch, c.cap, q);
    if (active) {
      if (g == 0) {
        st |= (unsigned)sm[L::FLAG];
        if (a.dist_out) a.dist_out[p] = sm[L::DIST];
      }
      if (a.vvd_out) {
        for (int k = g; k < N; k += G) a.vvd_out[p * N + k] = sm[L::VS + k];
      }
    }
    __syncwarp();
  }
  st = __reduce_or_sync(0xffffffffu, st);
  if (st != 0 && lane == 0 && a.status != nullptr) atomicOr(a.status, st);
}

template <int N, int MODE>
__global__ void __launch_bounds__(32) coop_backward_kernel(const PairArgs a, const ChunkArgs c) {
  typedef coop::Layout<N> L;
  constexpr int G = CoopCfg<N>::G;
  constexpr int PW = CoopCfg<N>::PW;
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int slot = lane / G;
  const int g = lane - slot * G;
  double* sm = smem + (slot < PW ? slot : 0) * L::kDoubles;
  double loss_acc = 0.0, gscale_acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * PW;
  for (int64_t base = (int64_t)blockIdx.x * PW; base < c.count; base += stride) {
    const int64_t q = base + slot;
    const int64_t p = c.base + q;
    bool active = slot < PW && q < c.count;
    const double* p1 = nullptr;
    const double* p2 = nullptr;
    int64_t i1 = 0, i2 = 0;
    if (active && !locate_pair<N>(a, p, &p1, &p2, &i1, &i2)) active = false;
    WarpExec ex{g, active};
    coop::split_backward<N>(ex, sm, p1, p2, c.scratch, c.cap, q);
    if (active) {
      const double dist = (MODE == kModeStep) ? a.dist_out[p] : 0.0;
      if (MODE == kModeStep && a.grad_wsum_w != nullptr && a.metric == kWsum) {
        const double gd = __ldg(a.graph_dist + p);
        const double r = a.scale * dist / gd;
        const double er = r * r - 1.0;
        const double dl_dd = (er > 0.0 ? 1.0 : (er < 0.0 ? -1.0 : 0.0)) * 2.0 * r * a.scale / gd;
        for (int k = g; k < N; k += G)
          if (a.wsum_w[k] > 0.0) atomicAdd(a.grad_wsum_w + k, dl_dd * a.vvd_out[p * N + k]);
      }
      emit_gradients<N, kUpper, MODE>(a, sm, g, p, i1, i2, dist, &loss_acc, &gscale_acc);
    }
    __syncwarp();
  }
  if (MODE == kModeStep) {
    loss_acc = warp_sum(loss_acc);
    gscale_acc = warp_sum(gscale_acc);
    if (lane == 0) {
      if (a.loss_out) atomicAdd(a.loss_out, loss_acc);
      if (a.grad_scale) atomicAdd(a.grad_scale, gscale_acc);
    }
  }
}

template <int N>
struct SplitCfg {
  static constexpr int PW = CoopCfg<N>::PW;
  static constexpr int kSmemFull = CoopCfg<N>::kSmemBytes;
  static constexpr int kSmemJ = PW * coop::LayoutJ<N>::kDoubles * (int)sizeof(double);
};

// MODE fwd: prologue + spectrum; fwd+save / step: + backward.  Step needs dist_out and (for wsum)
// vvd_out as per-pair temporaries: the ABI layer provides them inside the scratch allocation.
template <int N, int MODE>
static int launch_split(const PairArgs& a, double* scratch, int64_t cap, cudaStream_t s) {
  constexpr int PW = SplitCfg<N>::PW;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(coop_prologue_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SplitCfg<N>::kSmemFull);
    cudaFuncSetAttribute(coop_backward_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         SplitCfg<N>::kSmemFull);
    configured = true;
  }
  for (int64_t base = 0; base < a.num_pairs; base += cap) {
    ChunkArgs c;
    c.scratch = scratch;
    c.cap = cap;
    c.base = base;
    c.count = (a.num_pairs - base) < cap ? (a.num_pairs - base) : cap;
    const int64_t groups = (c.count + PW - 1) / PW;
    const int grid = grid_for(groups, 1, 32 * 8);
    coop_prologue_kernel<N><<<grid, 32, SplitCfg<N>::kSmemFull, s>>>(a, c);
    if (MODE == kModeFwd)
      coop_spectrum_kernel<N, false><<<grid, 32, SplitCfg<N>::kSmemJ, s>>>(a, c);
    else
      coop_spectrum_kernel<N, true><<<grid, 32, SplitCfg<N>::kSmemJ, s>>>(a, c);
    if (MODE != kModeFwd) coop_backward_kernel<N, MODE><<<grid, 32, SplitCfg<N>::kSmemFull, s>>>(a, c);
    const int rc = check_launch();
    if (rc) return rc;
  }
  return 0;
}

template <int N, int KIND, int MODE>
static int launch_coop(const PairArgs& a, cudaStream_t s) {
  constexpr int PW = CoopCfg<N>::PW;
#ifndef SY_COOP_SMEM_PAD
#define SY_COOP_SMEM_PAD 0   // experiment: unused shared memory per CTA, to lower the number of resident CTAs
#endif
  constexpr int smem = PW * CoopTraits<N, KIND>::kDoubles * (int)sizeof(double) + SY_COOP_SMEM_PAD;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(coop_kernel<N, KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    configured = true;
  }
  const int64_t groups = (a.num_pairs + PW - 1) / PW;
  const int grid = grid_for(groups, 1, 32 * 8);  // one warp per CTA; up to 8 waves of 32 CTAs per SM
  coop_kernel<N, KIND, MODE><<<grid, 32, smem, s>>>(a);
  return check_launch();
}

#endif  // SYMPA_PAIR_KERNELS_IMPL

}  // namespace sympa
