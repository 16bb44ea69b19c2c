// Warp-cooperative per-pair mathematics for the larger matrix sizes (n >= 5).
//
// Same algorithm as pair_math_impl.inc (see oracle/kernel_model.py) but a pair is worked on by a
// GROUP of G = ceil(n/2) lanes of one warp with every matrix resident in shared memory:
//   * dense stages: lane g owns output columns g and g+G, accumulates them in registers while it
//     streams the left operand from shared memory (all lanes of a group read the same address ->
//     broadcast), then writes its columns back;
//   * one-sided Jacobi: round-robin tournament, n/2 disjoint column pairs per round, one pair per
//     lane, columns pulled into registers, rotated, written back;
//   * backward without accumulating V: v_k = W^H g_k / sigma_k^2, so
//       G_W = sum_k (dL/dsigma_k / sigma_k^3) g_k g_k^H W = F W   (F Hermitian)
//     (halves the Jacobi work and the shared-memory footprint; the price is a relative error of
//     ~eps * cond(W) on the gradient, irrelevant against the reference's own 1e-10..1e-6 noise).
//
// The code is written as a sequence of STAGES.  Inside a stage a lane only reads buffers that no
// lane writes in that stage, so a stage boundary is the only synchronisation point: __syncwarp() on
// the device, nothing on the host, where tests/hostcheck runs the lanes of a group one after the
// other (class HostExec) - the very same templates are therefore checked on the CPU.
// State that lives across stages is in shared memory only.
#pragma once
#include "pair_math.cuh"

namespace sympa {
namespace coop {

// leading dimension of the shared-memory matrices: n + 1 for even n (odd: conflict-free column
// walks).  The unpadded n fits one more CTA per SM at n = 10 (6 instead of 5) but measured the same
// on the B200 (23.0 vs 23.5 ms per 2^20 pairs at n = 10, 11.2 vs 11.0 ms at n = 8): the cooperative
// kernel is bound by its serial stage structure, not by occupancy.
#ifndef SY_COOP_LD
#define SY_COOP_LD(N) (((N) % 2 == 0) ? (N) + 1 : (N))
#endif

template <int N, int NBUF>
struct LayoutT {
  static constexpr int NN = N * N;
  static constexpr int LD = SY_COOP_LD(N);
  static constexpr int BUF = N * LD;                    // one n x n buffer
  static constexpr int NP = (N % 2 == 0) ? N : N + 1;  // padded column count for the tournament
  static constexpr int G = NP / 2;                      // lanes per pair
  // small per-pair state after the NBUF matrix buffers
  static constexpr int SIG = NBUF * BUF, COEF = SIG + N, VS = COEF + N, RD = VS + N, VK = RD + N, DV = VK + N, RK = DV + N;
  static constexpr int WORST = RK + N;  // 2 * G convergence flags (double-buffered by sweep parity)
  static constexpr int DIST = WORST + 2 * G, FLAG = DIST + 1;
  static constexpr int kRaw = FLAG + 1;
  static constexpr int kDoubles = (kRaw % 2 == 0) ? kRaw + 1 : kRaw;  // odd stride between pair slots
};

// One-directional ring ordering of the register-resident Jacobi (WarpExec::jacobi): G lanes hold two columns each;
// after round r lane g sends one of them to lane (g + 1) mod G - bit g of mask r: 1 = the column it received after
// the previous round, 0 = the one it kept - and keeps the other.  The masks below were found by exhaustive search
// (every one of the N (N - 1) / 2 column pairs meets exactly once in the N - 1 = 2 G - 1 rounds of a sweep; the
// schedule is positional, so it holds from any arrangement and the last mask, 0, starts the next sweep on pairs that
// did not meet in the round before).  Packed: bit r * G + g.
template <int G>
SY_HD constexpr unsigned long long ring_send_masks() {
  static_assert(G >= 1 && G <= 5, "schedules are tabulated for 2..10 columns");
  constexpr unsigned m2[3] = {0, 1, 0};
  constexpr unsigned m3[5] = {0, 3, 3, 6, 0};
  constexpr unsigned m4[7] = {0, 7, 7, 14, 14, 13, 0};
  constexpr unsigned m5[9] = {0, 15, 15, 30, 30, 29, 29, 27, 0};
  unsigned long long packed = 0ull;
  for (int r = 0; r < 2 * G - 1; ++r) {
    const unsigned m = G == 2 ? m2[r < 3 ? r : 0] : (G == 3 ? m3[r < 5 ? r : 0] : (G == 4 ? m4[r < 7 ? r : 0] : (G == 5 ? m5[r < 9 ? r : 0] : 0u)));
    packed |= (unsigned long long)m << (r * G);
  }
  return packed;
}

// full pipeline: LI + six general buffers A0..A5
template <int N>
struct Layout : LayoutT<N, 7> {
  typedef LayoutT<N, 7> B;
  static constexpr int LI = 0, A0 = B::BUF, A1 = 2 * B::BUF, A2 = 3 * B::BUF, A3 = 4 * B::BUF, A4 = 5 * B::BUF,
                       A5 = 6 * B::BUF;
  // where the prologue leaves P and Q, where the Jacobi works, where upper_pair leaves its results
  static constexpr int PBUF = A0, QBUF = A4, GR = A1, GI = A2;
  static constexpr int GX2 = A4, GY2 = A3, GY1 = A0;
};

// Jacobi-only kernel (split path): just the two halves of G
template <int N>
struct LayoutJ : LayoutT<N, 2> {
  static constexpr int GR = 0, GI = LayoutT<N, 2>::BUF;
};

#if defined(__CUDA_ARCH__)
#define SY_LDG(p) __ldg(p)
#else
#define SY_LDG(p) (*(p))
#endif

// A stage: on the device the body runs once for this thread's lane g (if the lane is active) and
// ends with __syncwarp(); on the host it runs for g = 0 .. G-1 in turn.
#define SY_STAGE_BEGIN(ex) for (int g = (ex).first(); g < (ex).last(); ++g) {
#define SY_STAGE_END(ex) \
  }                      \
  (ex).sync();

struct HostExec {
  int G;
  int first() const { return 0; }
  int last() const { return G; }
  void sync() const {}
  bool any(bool p) const { return p; }
  // Jacobi on the column-major matrix (gr, gi), column stride LD; defined after jacobi_ring_host
  template <int N, bool IS_REAL>
  int jacobi(double* gr, double* gi, double* conv) const;
};

#ifndef SY_MM_UNROLL
#define SY_MM_UNROLL 2
#endif

// ---------------------------------------------------------------------------------------------
// All shared-memory matrices are n x n with leading dimension LD = Layout<N>::LD.
// dense helper: accumulate the two owned columns j0 = g, j1 = g + G (j1 may be >= N: ignored)
//   acc0[i] += sum_k A(i,k) * B(k, j0)      A(i,k) = TA ? a[k*LD + i] : a[i*LD + k]
//                                           B(k,j) = TB ? b[j*LD + k] : b[k*LD + j]
template <int N, bool TA, bool TB>
SY_HD void mm_cols(const double* a, const double* b, int j0, int j1, double* acc0, double* acc1) {
  constexpr int LD = Layout<N>::LD;
  const bool two = j1 < N;
  const int jj1 = two ? j1 : j0;
  constexpr int kMmUnroll = SY_MM_UNROLL;
#pragma unroll kMmUnroll
  for (int k = 0; k < N; ++k) {
    const double b0 = TB ? b[j0 * LD + k] : b[k * LD + j0];
    double b1 = TB ? b[jj1 * LD + k] : b[k * LD + jj1];
    b1 = two ? b1 : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double av = TA ? a[k * LD + i] : a[i * LD + k];
      acc0[i] += av * b0;
      acc1[i] += av * b1;
    }
  }
}

// two products that share the left operand: (A op) Bx and (A op) By for the owned columns; the column
// of A is read from shared memory once per k for four accumulator columns
template <int N, bool TA>
SY_HD void mm_cols2(const double* a, const double* bx, const double* by, int j0, int j1, double* x0, double* x1,
                    double* y0, double* y1) {
  constexpr int LD = Layout<N>::LD;
  const bool two = j1 < N;
  const int jj1 = two ? j1 : j0;
  constexpr int kMmUnroll = SY_MM_UNROLL;
#pragma unroll kMmUnroll
  for (int k = 0; k < N; ++k) {
    const double bx0 = bx[k * LD + j0], by0 = by[k * LD + j0];
    double bx1 = bx[k * LD + jj1], by1 = by[k * LD + jj1];
    bx1 = two ? bx1 : 0.0;
    by1 = two ? by1 : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double av = TA ? a[k * LD + i] : a[i * LD + k];
      x0[i] += av * bx0;
      x1[i] += av * bx1;
      y0[i] += av * by0;
      y1[i] += av * by1;
    }
  }
}

template <int N>
SY_HD void zero2(double* a, double* b) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a[i] = 0.0;
    b[i] = 0.0;
  }
}

// write owned columns: c[i*LD + j] = s * acc[i] (+ diag on the diagonal)
template <int N>
SY_HD void put_cols(double* c, int j0, int j1, const double* acc0, const double* acc1, double s, double diag) {
  constexpr int LD = Layout<N>::LD;
#pragma unroll
  for (int i = 0; i < N; ++i) c[i * LD + j0] = s * acc0[i] + ((i == j0) ? diag : 0.0);
  if (j1 < N) {
#pragma unroll
    for (int i = 0; i < N; ++i) c[i * LD + j1] = s * acc1[i] + ((i == j1) ? diag : 0.0);
  }
}

// c = s * (A op B) (+ diag I) for the owned columns - one stage body
template <int N, bool TA, bool TB>
SY_HD void mm_stage(const double* a, const double* b, double* c, int g, double s, double diag) {
  constexpr int G = Layout<N>::G;
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, TA, TB>(a, b, g, g + G, a0, a1);
  put_cols<N>(c, g, g + G, a0, a1, s, diag);
}

// Inverse Cholesky factor, cooperative: a (full symmetric, lower part used) -> li (full, upper part
// zeroed); l = scratch buffer, rd = n doubles.  N + 1 stages.
template <int N, class Ex>
SY_HD void chol_inv(Ex& ex, const double* a, double* l, double* li, double* rd, double* flag,
                    bool keep_l = false) {  // keep_l: leave a clean lower-triangular L in `l` (upper part zeroed)
  constexpr int G = Layout<N>::G;
  constexpr int LD = Layout<N>::LD;
  // The k-loops below run over a fixed, fully unrolled range with a predicate instead of a runtime
  // trip count: all the shared-memory loads of a dot product are then issued back to back and only
  // the FMA chain stays serial (these stages are latency-bound, one warp per scheduler).
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    SY_STAGE_BEGIN(ex)
    double lj[N];
#pragma unroll
    for (int k = 0; k < N; ++k) lj[k] = (k < j) ? l[j * LD + k] : 0.0;
    double d0 = a[j * LD + j], d1 = 0.0;
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      d0 -= lj[k] * lj[k];
      if (k + 1 < N) d1 -= lj[k + 1] * lj[k + 1];
    }
    const double d = d0 + d1;
    const double r = sy_rsqrt(d);
    for (int i = j + g; i < N; i += G) {
      if (i == j) {
        l[j * LD + j] = d * r;
        rd[j] = r;
        if (!(d > 0.0)) *flag = 1.0;
      } else {
        double lik[N];
#pragma unroll
        for (int k = 0; k < N; ++k) lik[k] = (k < j) ? l[i * LD + k] : 0.0;
        double s0 = a[i * LD + j], s1 = 0.0;
#pragma unroll
        for (int k = 0; k < N; k += 2) {
          s0 -= lik[k] * lj[k];
          if (k + 1 < N) s1 -= lik[k + 1] * lj[k + 1];
        }
        l[i * LD + j] = (s0 + s1) * r;
      }
    }
    SY_STAGE_END(ex)
  }
  // inverse of the triangular factor: lane g solves for columns g and g + G, both at once
  SY_STAGE_BEGIN(ex)
  const int c0 = g, c1 = g + G;
  const bool two = c1 < N;
  double x0[N], x1[N];  // the two columns of li being built (entries above the diagonal are zero)
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double li_row[N];
#pragma unroll
    for (int k = 0; k < N; ++k) li_row[k] = (k < i) ? l[i * LD + k] : 0.0;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (k < i) {
        s0 += li_row[k] * x0[k];
        s1 += li_row[k] * x1[k];
      }
    }
    const double r = rd[i];
    x0[i] = (i == c0) ? r : ((i > c0) ? -s0 * r : 0.0);
    x1[i] = (two && i == c1) ? r : ((two && i > c1) ? -s1 * r : 0.0);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    li[i * LD + c0] = x0[i];
    if (two) li[i * LD + c1] = x1[i];
  }
  if (keep_l) {
    for (int i = 0; i < c0; ++i) l[i * LD + c0] = 0.0;
    if (two)
      for (int i = 0; i < c1; ++i) l[i * LD + c1] = 0.0;
  }
  SY_STAGE_END(ex)
}

// (E + iF)^-1 = U + iV, E SPD (buffer e), F symmetric (buffer f).  Scratch: s0, s1, s2.
// e, f, u, s0, s1, s2 are six distinct buffers; v may alias f (F is dead once T is formed).
//   lei -> s0 ; T = lei F -> s1 ; S = E + T^T T -> s2 ; lsi -> e (E dead) ; U = lsi^T lsi -> u ;
//   TU -> s2 ; V = -lei^T TU -> v
template <int N, class Ex>
SY_HD void inv_spd_real(Ex& ex, double* e, const double* f, double* u, double* v, double* s0, double* s1, double* s2,
                        double* rd, double* flag) {
  constexpr int G = Layout<N>::G;
  constexpr int LD = Layout<N>::LD;
  chol_inv<N>(ex, e, s1, s0, rd, flag);  // s1 = scratch L, s0 = lei
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(s0, f, s1, g, 1.0, 0.0);  // T = lei F
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(s1, s1, g, g + G, a0, a1);  // T^T T
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s2[i * LD + g] = a0[i] + e[i * LD + g];
    if (g + G < N) s2[i * LD + g + G] = a1[i] + e[i * LD + g + G];
  }
  SY_STAGE_END(ex)
  chol_inv<N>(ex, s2, u, e, rd, flag);  // u = scratch L, e = lsi
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(e, e, u, g, 1.0, 0.0);  // U = lsi^T lsi
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(s1, u, s2, g, 1.0, 0.0);  // TU
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(s0, s2, v, g, -1.0, 0.0);  // V = -lei^T TU
  SY_STAGE_END(ex)
}

// tournament pairing: round r in [0, NP-1), slot k in [0, NP/2)
template <int N>
SY_HD void tournament(int r, int k, int* p, int* q) {
  constexpr int NP = Layout<N>::NP;
  (void)NP;
  loc::tournament_pair<N>(r, k, p, q);
}

// Orthogonalise one column pair held in registers (complex, or real when IS_REAL).  Returns the
// convergence flags of the pair (kConvLoose | kConvStrict | kConvClose, see jacobi_rotation).
template <int N, bool IS_REAL>
SY_HD double column_norm2(const double* pr, const double* pi) {
  double a0 = 0.0, a1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a0 += pr[i] * pr[i];
    if (!IS_REAL) a1 += pi[i] * pi[i];
  }
  return a0 + a1;
}

// *pn, *qn: the squared norms of the two columns, carried along with them (recomputed by the caller
// at the start of every sweep, updated here: see reg::jacobi_svd).
template <int N, bool IS_REAL>
SY_HD unsigned rotate_columns(double* pr, double* pi, double* qr, double* qi, double* pn, double* qn) {
  // two partial sums per quantity: shorter dependency chains
  double cr0 = 0.0, ci0 = 0.0, cr1 = 0.0, ci1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    cr0 += pr[i] * qr[i];
    if (!IS_REAL) {
      cr1 += pi[i] * qi[i];
      ci0 += pr[i] * qi[i];
      ci1 += pi[i] * qr[i];
    }
  }
  unsigned conv = 0u;
  double c, sr, si, xf;
  // unconditional: a skipped rotation comes back as the identity (c = 1, s = 0, xf = 0) - no divergence between
  // the groups of a warp in the last sweeps
  loc::jacobi_rotation_t<loc::jacobi_stop_log2<N>()>(*pn, *qn, cr0 + cr1, ci0 - ci1, &c, &sr, &si, &xf, &conv);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double a0 = pr[i], b0 = qr[i];
    if (IS_REAL) {
      pr[i] = fma(c, a0, -(sr * b0));
      qr[i] = fma(c, b0, sr * a0);
    } else {
      const double a1 = pi[i], b1 = qi[i];
      pr[i] = fma(c, a0, -fma(si, b1, sr * b0));
      pi[i] = fma(c, a1, -fma(-si, b0, sr * b1));
      qr[i] = fma(c, b0, fma(-si, a1, sr * a0));
      qi[i] = fma(c, b1, fma(si, a0, sr * a1));
    }
  }
  *pn -= xf;
  *qn += xf;
  return conv;
}

// Register-resident one-sided Jacobi in the one-directional ring ordering (ring_send_masks): lane g of a group
// holds a "top" and a "bottom" column; every round each lane rotates its own pair, then passes one of the two to
// its right neighbour, which visits every pair once per NP - 1 rounds.  The columns never touch shared memory between the
// initial load and the final store (the shared-memory version below moved both columns of every
// pair through shared memory every round and was bound by bank-conflicted LDS/STS wavefronts).
// An odd n gets a zero padding column, which no rotation ever changes.  Columns carry their index so
// that they are stored back in place.  This is the host emulation (all lanes of the group in one
// call); WarpExec implements the same schedule with warp shuffles.
template <int N, bool IS_REAL>
inline int jacobi_ring_host(int G, double* gr, double* gi) {
  typedef LayoutT<N, 2> L;
  constexpr int NP = L::NP, LD = L::LD;
  double tr[L::G][N], ti[L::G][N], br[L::G][N], bi[L::G][N];
  int tid[L::G], bid[L::G];
  for (int g = 0; g < G; ++g) {
    tid[g] = 2 * g;
    bid[g] = 2 * g + 1;
    for (int i = 0; i < N; ++i) {
      tr[g][i] = gr[tid[g] * LD + i];
      ti[g][i] = IS_REAL ? 0.0 : gi[tid[g] * LD + i];
      br[g][i] = bid[g] < N ? gr[bid[g] * LD + i] : 0.0;
      bi[g][i] = (bid[g] < N && !IS_REAL) ? gi[bid[g] * LD + i] : 0.0;
    }
  }
  int sweep = 0;
  double tn[L::G], bn[L::G];  // carried squared column norms
  for (; sweep < kMaxSweeps; ++sweep) {
    unsigned conv = 0u;
    for (int g = 0; g < G; ++g) {
      tn[g] = column_norm2<N, IS_REAL>(tr[g], ti[g]);
      bn[g] = column_norm2<N, IS_REAL>(br[g], bi[g]);
    }
    for (int r = 0; r < NP - 1; ++r) {
      for (int g = 0; g < G; ++g) conv |= rotate_columns<N, IS_REAL>(tr[g], ti[g], br[g], bi[g], &tn[g], &bn[g]);
      if (G > 1) {
        // one-directional ring: lane g passes its bottom column (bit 1: the one received last) or its top column
        // (bit 0) to lane g + 1 and keeps the other as its top - the schedule of WarpExec::jacobi
        const unsigned long long sched = ring_send_masks<L::G>();
        double sr_[L::G][N], si_[L::G][N], sn_[L::G];
        int sid_[L::G];
        for (int g = 0; g < G; ++g) {
          const bool pass_b = ((sched >> (r * L::G + g)) & 1ull) != 0ull;
          for (int i = 0; i < N; ++i) {
            sr_[g][i] = pass_b ? br[g][i] : tr[g][i];
            si_[g][i] = pass_b ? bi[g][i] : ti[g][i];
            if (!pass_b) {
              tr[g][i] = br[g][i];
              ti[g][i] = bi[g][i];
            }
          }
          sn_[g] = pass_b ? bn[g] : tn[g];
          sid_[g] = pass_b ? bid[g] : tid[g];
          if (!pass_b) {
            tn[g] = bn[g];
            tid[g] = bid[g];
          }
        }
        for (int g = 0; g < G; ++g) {
          const int src = g == 0 ? G - 1 : g - 1;
          for (int i = 0; i < N; ++i) {
            br[g][i] = sr_[src][i];
            bi[g][i] = si_[src][i];
          }
          bn[g] = sn_[src];
          bid[g] = sid_[src];
        }
      }
    }
    if (loc::jacobi_sweep_is_last(conv)) {
      ++sweep;
      break;
    }
  }
  for (int g = 0; g < G; ++g) {
    for (int i = 0; i < N; ++i) {
      if (tid[g] < N) gr[tid[g] * LD + i] = tr[g][i];
      if (bid[g] < N) gr[bid[g] * LD + i] = br[g][i];
      if (!IS_REAL) {
        if (tid[g] < N) gi[tid[g] * LD + i] = ti[g][i];
        if (bid[g] < N) gi[bid[g] * LD + i] = bi[g][i];
      }
    }
  }
  return sweep;
}

// One-sided Jacobi on the column-major complex matrix (gr, gi) in shared memory, column stride LD
// (no V).  `conv` holds 2 * G not-converged flags, double-buffered by sweep parity.
template <int N, bool IS_REAL>
int HostExec::jacobi(double* gr, double* gi, double*) const {
  return jacobi_ring_host<N, IS_REAL>(G, gr, gi);
}

// The first implementation, kept for reference: same mathematics with the columns living in shared
// memory and a tournament pairing (one pair per lane per round, __syncwarp between rounds).
template <int N, bool IS_REAL, class Ex>
SY_HD int jacobi_smem(Ex& ex, double* gr, double* gi, double* conv) {
  constexpr int NP = Layout<N>::NP;
  constexpr int G = Layout<N>::G;
  constexpr int LD = Layout<N>::LD;
  int sweep = 0;
#pragma unroll 1
  for (; sweep < kMaxSweeps; ++sweep) {
    double* wbuf = conv + (sweep & 1) * G;
#pragma unroll 1
    for (int r = 0; r < NP - 1; ++r) {
      SY_STAGE_BEGIN(ex)
      int p, q;
      tournament<N>(r, g, &p, &q);
      double notconv = 0.0;
      if (q < N) {  // q == N is the padding column of an odd n
        double pr[N], pi[N], qr[N], qi[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          pr[i] = gr[p * LD + i];
          qr[i] = gr[q * LD + i];
          pi[i] = IS_REAL ? 0.0 : gi[p * LD + i];
          qi[i] = IS_REAL ? 0.0 : gi[q * LD + i];
        }
        // two partial sums per quantity: shorter dependency chains
        double al0 = 0.0, be0 = 0.0, cr0 = 0.0, ci0 = 0.0, al1 = 0.0, be1 = 0.0, cr1 = 0.0, ci1 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          al0 += pr[i] * pr[i];
          al1 += pi[i] * pi[i];
          be0 += qr[i] * qr[i];
          be1 += qi[i] * qi[i];
          cr0 += pr[i] * qr[i];
          cr1 += pi[i] * qi[i];
          ci0 += pr[i] * qi[i];
          ci1 += pi[i] * qr[i];
        }
        const double al = al0 + al1, be = be0 + be1, cr = cr0 + cr1, ci = ci0 - ci1;
        unsigned conv = 0u;
        double c, sr, si, xf;
        if (loc::jacobi_rotation(al, be, cr, ci, loc::jacobi_stop_ratio2<N>(), &c, &sr, &si, &xf, &conv)) {
#pragma unroll
          for (int i = 0; i < N; ++i) {
            gr[p * LD + i] = c * pr[i] - (sr * qr[i] + si * qi[i]);
            gr[q * LD + i] = c * qr[i] + (sr * pr[i] - si * pi[i]);
            if (!IS_REAL) {
              gi[p * LD + i] = c * pi[i] - (sr * qi[i] - si * qr[i]);
              gi[q * LD + i] = c * qi[i] + (sr * pi[i] + si * pr[i]);
            }
          }
        }
        notconv = (conv & loc::kConvLoose) ? 1.0 : 0.0;
      }
      wbuf[g] = (r == 0) ? notconv : (notconv > wbuf[g] ? notconv : wbuf[g]);
      SY_STAGE_END(ex)
    }
    double w = 0.0;
    for (int k = 0; k < G; ++k) w = wbuf[k] > w ? wbuf[k] : w;
    const bool more = !(w < 0.5);
    if (!ex.any(more)) {
      ++sweep;
      break;
    }
  }
  return sweep;
}

// ---------------------------------------------------------------------------------------------
// Metric tail, lanes cooperate: sm[SIG..] holds the (unsorted) singular values.
//  stage 1: lane k ranks its sigma, computes v_k and dv_k/dsigma_k;
//  stage 2: every lane sums the n values of v (cheap, identical on all lanes), lane k writes its
//           backward coefficient (dL/dsigma_k) / sigma_k^3; lane 0 writes dist and the status.
template <int N, class L, class Ex>
SY_HD void metric_tail(Ex& ex, double* sm, int metric, const double* wsum_w, int sweeps) {
  constexpr int G = L::G;
  SY_STAGE_BEGIN(ex)
  for (int k = g; k < N; k += G) {
    const double d = sm[L::SIG + k];
    int r = 0;
    for (int j = 0; j < N; ++j) {
      const double o = sm[L::SIG + j];
      r += (o < d || (o == d && j < k)) ? 1 : 0;
    }
    const double om = 1.0 - d;
    const bool clamped = om < kEpsF64;
    const double v = clamped ? log((1.0 + d) / kEpsF64) : log1p(2.0 * d / om);   // siegel_manifold.py:69-70
    const double dv = clamped ? 1.0 / (1.0 + d) : 2.0 / (om * (1.0 + d));
    sm[L::VK + k] = v;
    sm[L::DV + k] = dv;
    sm[L::RK + k] = (double)r;
    sm[L::VS + r] = v;
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double dist = 0.0;
  bool above = false;
  if (metric == kRiem) {
    for (int k = 0; k < N; ++k) dist += sm[L::VK + k] * sm[L::VK + k];
    dist = sqrt(dist);
  }
  for (int k = 0; k < N; ++k) above = above || (sm[L::SIG + k] > 1.01);
  double part = 0.0;  // this lane's share of a linear metric
  for (int k = g; k < N; k += G) {
    const double sig = sm[L::SIG + k];
    const double v = sm[L::VK + k];
    const int r = (int)sm[L::RK + k];
    double gv;
    if (metric == kRiem) {
      gv = dist > 0.0 ? v / dist : 0.0;
    } else {
      if (metric == kFone) gv = 1.0;
      else if (metric == kFinf) gv = (r == N - 1) ? 1.0 : 0.0;
      else if (metric == kFmin) gv = 2.0 * (double)r;
      else { gv = wsum_w[r]; gv = gv > 0.0 ? gv : 0.0; }
      part += gv * v;
    }
    // dL/dsigma_k / sigma_k^3   (u_k = g_k / sigma_k, v_k = W^H g_k / sigma_k^2)
    sm[L::COEF + k] = sig > 1e-150 ? gv * sm[L::DV + k] / (sig * sig * sig) : 0.0;
  }
  if (metric != kRiem) sm[L::WORST + g] = part;  // the convergence flags are dead by now
  if (g == 0) {
    unsigned st = (sm[L::FLAG] != 0.0) ? kStatusNotPD : 0u;
    if (sweeps >= kMaxSweeps) st |= kStatusNoConverge;
    if (above) st |= kStatusTakagiAboveOne;
    sm[L::FLAG] = (double)st;
    if (metric == kRiem) sm[L::DIST] = dist;
  }
  SY_STAGE_END(ex)
  if (metric != kRiem) {
    SY_STAGE_BEGIN(ex)
    if (g == 0) {
      double dist = 0.0;
      for (int k = 0; k < G; ++k) dist += sm[L::WORST + k];
      sm[L::DIST] = dist;
    }
    SY_STAGE_END(ex)
  }
  SY_STAGE_BEGIN(ex)
  if (g == 0) {
    const double dist = sm[L::DIST];
    if (!(dist == dist) || dist > 1e300) sm[L::FLAG] = (double)((unsigned)sm[L::FLAG] | kStatusNonFinite);
  }
  SY_STAGE_END(ex)
}

// ---------------------------------------------------------------------------------------------
// Upper half space, one pair, cooperative, in three parts (fused by upper_pair, or run as three
// kernels with the state parked in global scratch - see coop_kernels.cuh):
//   upper_prologue : gather -> LI = chol(Y1)^-1, (A + iC)^-1 = P - iQ        (sm[LI], sm[PBUF], sm[QBUF])
//   w_from_pq + jacobi + sigma + metric_tail : W -> Takagi values -> vvd, dist, backward coefficients
//   upper_backward : unit gradients                                          (sm[GX2], sm[GY2], sm[GY1])
// p1 / p2 point at the (2, n, n) rows in global memory.
// Where the upper-half operands D = X2 - X1, Y1, Y2 come from: straight from the two (2, n, n) rows
// in global memory (symmetrised on the fly) ...
template <int N>
struct GlobalIn {
  static constexpr bool kResetFlag = true;
  const double* p1;
  const double* p2;
  SY_HD double d(int e, int et, int) const {
    return 0.5 * ((SY_LDG(p2 + e) + SY_LDG(p2 + et)) - (SY_LDG(p1 + e) + SY_LDG(p1 + et)));
  }
  SY_HD double y1(int e, int et, int, bool) const { return 0.5 * (SY_LDG(p1 + N * N + e) + SY_LDG(p1 + N * N + et)); }
  SY_HD double y2(int e, int et, int, bool) const { return 0.5 * (SY_LDG(p2 + N * N + e) + SY_LDG(p2 + N * N + et)); }
};
// ... or, for the bounded domain, from the inverse Cayley transforms kept in shared memory:
//   Z_k = i (2 N_k - I),  N_k = (I - z_k)^-1 = U_k + i V_k   ->   X_k = -2 V_k,  Y_k = 2 U_k - I
template <int N>
struct CayleyIn {
  static constexpr bool kResetFlag = false;  // the flag already carries the status of the Cayley inverses
  const double* u1;
  const double* v1;
  const double* u2;
  const double* v2;
  SY_HD double d(int, int, int s) const { return -2.0 * (v2[s] - v1[s]); }
  SY_HD double y1(int, int, int s, bool diag) const { return 2.0 * u1[s] - (diag ? 1.0 : 0.0); }
  SY_HD double y2(int, int, int s, bool diag) const { return 2.0 * u2[s] - (diag ? 1.0 : 0.0); }
};

template <int N, class In, class Ex>
SY_HD void upper_prologue(Ex& ex, double* sm, const In& in) {
  typedef Layout<N> L;
  constexpr int G = L::G;
  constexpr int NN = L::NN;
  constexpr int LD = L::LD;
  double* li = sm + L::LI;
  double* a0 = sm + L::A0;
  double* a1 = sm + L::A1;
  double* a2 = sm + L::A2;
  double* a3 = sm + L::A3;
  double* a4 = sm + L::A4;
  double* a5 = sm + L::A5;
  double* rd = sm + L::RD;
  double* flag = sm + L::FLAG;
  (void)G; (void)NN; (void)LD; (void)li; (void)a0; (void)a1; (void)a2; (void)a3; (void)a4; (void)a5; (void)rd; (void)flag;
  // ---- load: a0 = D = X2 - X1, a1 = Y1, a2 = Y2 (symmetrised)
  SY_STAGE_BEGIN(ex)
  if (g == 0 && In::kResetFlag) *flag = 0.0;
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N, et = j * N + i, s = i * LD + j;
    a0[s] = in.d(e, et, s);
    a1[s] = in.y1(e, et, s, i == j);
    a2[s] = in.y2(e, et, s, i == j);
  }
  SY_STAGE_END(ex)

  chol_inv<N>(ex, a1, a3, li, rd, flag);  // li = chol(Y1)^-1, scratch a3

  SY_STAGE_BEGIN(ex)
  double x0[N], x1[N], y0[N], y1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  mm_cols2<N, false>(li, a0, a2, g, g + G, x0, x1, y0, y1);
  put_cols<N>(a1, g, g + G, x0, x1, 1.0, 0.0);  // a1 = li D
  put_cols<N>(a3, g, g + G, y0, y1, 1.0, 0.0);  // a3 = li Y2
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, true>(a1, li, a0, g, -1.0, 0.0);  // a0 = -A = -(li D) li^T
  mm_stage<N, false, true>(a3, li, a2, g, 1.0, 1.0);   // a2 = C = (li Y2) li^T + I
  SY_STAGE_END(ex)

  // ---- (C - iA)^-1 = Q + iP  ->  (A + iC)^-1 = P - iQ :   Q = a4, P = a0
  inv_spd_real<N>(ex, a2, a0, a4, a0, a5, a1, a3, rd, flag);

}

// W = (I - 2Q) - 2iP, column-major with column stride LD (P, Q row-major with leading dimension ldpq)
template <int N, class Ex>
SY_HD void w_from_pq(Ex& ex, const double* P, const double* Q, int ldpq, double* gr, double* gi) {
  typedef LayoutT<N, 2> L;
  constexpr int G = L::G;
  constexpr int LD = L::LD;
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < L::NN; e += G) {
    const int c = e / N, r = e - c * N;
    gr[c * LD + r] = ((r == c) ? 1.0 : 0.0) - 2.0 * Q[r * ldpq + c];  // generic loads: shared or global
    gi[c * LD + r] = -2.0 * P[r * ldpq + c];
  }
  SY_STAGE_END(ex)
}

// Jacobi + singular values + metric on the G buffers of layout LT (Layout<N> or LayoutJ<N>)
template <int N, class LT, class Ex>
SY_HD void spectrum(Ex& ex, double* sm, int metric, const double* wsum_w) {
  constexpr int G = LT::G;
  constexpr int LD = LT::LD;
  double* gr = sm + LT::GR;
  double* gi = sm + LT::GI;
  const int sweeps = ex.template jacobi<N, false>(gr, gi, sm + LT::WORST);
  SY_STAGE_BEGIN(ex)
  for (int k = g; k < N; k += G) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < N; ++i) {
      s0 += gr[k * LD + i] * gr[k * LD + i];
      s1 += gi[k * LD + i] * gi[k * LD + i];
    }
    sm[LT::SIG + k] = sqrt(s0 + s1);
  }
  SY_STAGE_END(ex)
  metric_tail<N, LT>(ex, sm, metric, wsum_w, sweeps);
}

// backward: needs sm[LI], P = sm[PBUF], Q = sm[QBUF], G = (sm[GR], sm[GI]) column-major, sm[COEF]
template <int N, class In, class Ex>
SY_HD void upper_backward(Ex& ex, double* sm, const In& in) {
  typedef Layout<N> L;
  constexpr int G = L::G;
  constexpr int NN = L::NN;
  constexpr int LD = L::LD;
  double* li = sm + L::LI;
  double* a0 = sm + L::A0;
  double* a1 = sm + L::A1;
  double* a2 = sm + L::A2;
  double* a3 = sm + L::A3;
  double* a4 = sm + L::A4;
  double* a5 = sm + L::A5;
  double* rd = sm + L::RD;
  double* flag = sm + L::FLAG;
  (void)G; (void)NN; (void)LD; (void)li; (void)a0; (void)a1; (void)a2; (void)a3; (void)a4; (void)a5; (void)rd; (void)flag;
  double* P = a0;
  double* Q = a4;
  // ---- F = sum_k coef_k g_k g_k^H  (Hermitian):  a3 = Re F, a5 = Im F
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  const bool two = j1 < N;
  const int jj1 = two ? j1 : j0;
  double fr0[N], fi0[N], fr1[N], fi1[N];
  zero2<N>(fr0, fi0);
  zero2<N>(fr1, fi1);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double c = sm[L::COEF + k];
    const double b0r = c * a1[k * LD + j0], b0i = c * a2[k * LD + j0];
    const double b1r = two ? c * a1[k * LD + jj1] : 0.0, b1i = two ? c * a2[k * LD + jj1] : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double ar = a1[k * LD + i], ai = a2[k * LD + i];
      fr0[i] += ar * b0r + ai * b0i;  // g_k[i] * conj(g_k[j])
      fi0[i] += ai * b0r - ar * b0i;
      fr1[i] += ar * b1r + ai * b1i;
      fi1[i] += ai * b1r - ar * b1i;
    }
  }
  put_cols<N>(a3, j0, j1, fr0, fr1, 1.0, 0.0);
  put_cols<N>(a5, j0, j1, fi0, fi1, 1.0, 0.0);
  SY_STAGE_END(ex)

  // ---- GW = F W = F - 2 F (Q + iP):  a1 = Re = Fr - 2 (Fr Q - Fi P),  a2 = Im = Fi - 2 (Fr P + Fi Q)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  mm_cols2<N, false>(a3, Q, P, j0, j1, x0, x1, y0, y1);  // Fr Q, Fr P
  double z0[N], z1[N], w0[N], w1[N];
  zero2<N>(z0, z1);
  zero2<N>(w0, w1);
  mm_cols2<N, false>(a5, P, Q, j0, j1, z0, z1, w0, w1);  // Fi P, Fi Q
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a1[i * LD + j0] = a3[i * LD + j0] - 2.0 * (x0[i] - z0[i]);
    a2[i * LD + j0] = a5[i * LD + j0] - 2.0 * (y0[i] + w0[i]);
    if (j1 < N) {
      a1[i * LD + j1] = a3[i * LD + j1] - 2.0 * (x1[i] - z1[i]);
      a2[i * LD + j1] = a5[i * LD + j1] - 2.0 * (y1[i] + w1[i]);
    }
  }
  SY_STAGE_END(ex)

  // ---- G_N = 2i sym(GW) = i (GW + GW^T):  a3 = Re = -(GWi + GWi^T),  a5 = Im = GWr + GWr^T
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N, s = i * LD + j, t = j * LD + i;
    a3[s] = -(a2[s] + a2[t]);
    a5[s] = a1[s] + a1[t];
  }
  SY_STAGE_END(ex)

  // ---- Tm = conj(N) G_N, conj(N) = P + iQ:  a1 = P GNr - Q GNi,  a2 = P GNi + Q GNr
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N], z0[N], z1[N], w0[N], w1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  zero2<N>(z0, z1);
  zero2<N>(w0, w1);
  mm_cols2<N, false>(P, a3, a5, j0, j1, x0, x1, z0, z1);  // P GNr, P GNi
  mm_cols2<N, false>(Q, a5, a3, j0, j1, y0, y1, w0, w1);  // Q GNi, Q GNr
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a1[i * LD + j0] = x0[i] - y0[i];
    a2[i * LD + j0] = z0[i] + w0[i];
    if (j1 < N) {
      a1[i * LD + j1] = x1[i] - y1[i];
      a2[i * LD + j1] = z1[i] + w1[i];
    }
  }
  SY_STAGE_END(ex)

  // ---- G_M = -Tm conj(N):  a3 = G_A = -(Tmr P - Tmi Q),  a5 = G_B = -(Tmr Q + Tmi P)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N], z0[N], z1[N], w0[N], w1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  zero2<N>(z0, z1);
  zero2<N>(w0, w1);
  mm_cols2<N, false>(a1, P, Q, j0, j1, x0, x1, z0, z1);  // Tmr P, Tmr Q
  mm_cols2<N, false>(a2, Q, P, j0, j1, y0, y1, w0, w1);  // Tmi Q, Tmi P
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a3[i * LD + j0] = -(x0[i] - y0[i]);
    a5[i * LD + j0] = -(z0[i] + w0[i]);
    if (j1 < N) {
      a3[i * LD + j1] = -(x1[i] - y1[i]);
      a5[i * LD + j1] = -(z1[i] + w1[i]);
    }
  }
  SY_STAGE_END(ex)

  // ---- reload D -> a0, Y2 -> a4 (P, Q are dead)
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N, et = j * N + i, s = i * LD + j;
    a0[s] = in.d(e, et, s);
    a4[s] = in.y2(e, et, s, i == j);
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double x0[N], x1[N], y0[N], y1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  mm_cols2<N, false>(li, a0, a4, g, g + G, x0, x1, y0, y1);
  put_cols<N>(a1, g, g + G, x0, x1, 1.0, 0.0);  // a1 = T1 = li D
  put_cols<N>(a2, g, g + G, y0, y1, 1.0, 0.0);  // a2 = T2 = li Y2
  SY_STAGE_END(ex)
  // ---- a0 = G_Li = 2 (G_A T1 + G_B T2)   (only its lower triangle is used)
  SY_STAGE_BEGIN(ex)
  double x0[N], x1[N];
  zero2<N>(x0, x1);
  mm_cols<N, false, false>(a3, a1, g, g + G, x0, x1);
  mm_cols<N, false, false>(a5, a2, g, g + G, x0, x1);
  put_cols<N>(a0, g, g + G, x0, x1, 2.0, 0.0);
  SY_STAGE_END(ex)
  // ---- a4 = K_s: 0.5 * tril(G_Li li^T) mirrored to a full symmetric matrix
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N;
    if (i >= j) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        s0 += (k <= j) ? a0[i * LD + k] * li[j * LD + k] : 0.0;
        if (k + 1 < N) s1 += (k + 1 <= j) ? a0[i * LD + k + 1] * li[j * LD + k + 1] : 0.0;
      }
      a4[i * LD + j] = 0.5 * (s0 + s1);
      a4[j * LD + i] = 0.5 * (s0 + s1);
    }
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(a4, li, a1, g, 1.0, 0.0);  // a1 = K_s li
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(li, a1, a0, g, -1.0, 0.0);  // a0 = G_Y1 = -li^T K_s li
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(a3, li, a1, g, 1.0, 0.0);  // a1 = G_A li
  mm_stage<N, false, false>(a5, li, a2, g, 1.0, 0.0);  // a2 = G_B li
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double x0[N], x1[N], y0[N], y1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  mm_cols2<N, true>(li, a1, a2, g, g + G, x0, x1, y0, y1);
  put_cols<N>(a4, g, g + G, x0, x1, 1.0, 0.0);  // a4 = G_X2 = li^T G_A li
  put_cols<N>(a3, g, g + G, y0, y1, 1.0, 0.0);  // a3 = G_Y2 = li^T G_B li
  SY_STAGE_END(ex)
}

// park / fetch one n x n buffer (leading dimension LD in shared memory, compact n*n in global scratch)
template <int N, class Ex>
SY_HD void park(Ex& ex, const double* smbuf, double* gbuf) {
  typedef LayoutT<N, 2> L;
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < L::NN; e += L::G) {
    const int i = e / N, j = e - i * N;
    gbuf[e] = smbuf[i * L::LD + j];
  }
  SY_STAGE_END(ex)
}
template <int N, class Ex>
SY_HD void fetch(Ex& ex, const double* gbuf, double* smbuf) {
  typedef LayoutT<N, 2> L;
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < L::NN; e += L::G) {
    const int i = e / N, j = e - i * N;
    smbuf[i * L::LD + j] = gbuf[e];
  }
  SY_STAGE_END(ex)
}

// scratch layout of the split path, per chunk of `cap` pairs (doubles): five n x n planes + coef
template <int N>
struct Scratch {
  static constexpr int64_t kPerPair = 5 * N * N + N;
  static SY_HD double* li(double* s, int64_t cap, int64_t p) { return s + (0 * cap + p) * N * N; }
  static SY_HD double* pb(double* s, int64_t cap, int64_t p) { return s + (1 * cap + p) * N * N; }
  static SY_HD double* qb(double* s, int64_t cap, int64_t p) { return s + (2 * cap + p) * N * N; }
  static SY_HD double* gr(double* s, int64_t cap, int64_t p) { return s + (3 * cap + p) * N * N; }
  static SY_HD double* gi(double* s, int64_t cap, int64_t p) { return s + (4 * cap + p) * N * N; }
  static SY_HD double* coef(double* s, int64_t cap, int64_t p) { return s + 5 * cap * N * N + p * N; }
};

// split path, part 1: prologue, then park LI, P, Q
template <int N, class Ex>
SY_HD void split_prologue(Ex& ex, double* sm, const double* p1, const double* p2, double* scratch, int64_t cap,
                          int64_t slot) {
  typedef Layout<N> L;
  const GlobalIn<N> in{p1, p2};
  upper_prologue<N>(ex, sm, in);
  park<N>(ex, sm + L::LI, Scratch<N>::li(scratch, cap, slot));
  park<N>(ex, sm + L::PBUF, Scratch<N>::pb(scratch, cap, slot));
  park<N>(ex, sm + L::QBUF, Scratch<N>::qb(scratch, cap, slot));
}
// part 2: W from the parked P, Q; Jacobi; metric; park G (column-major planes) and the coefficients
template <int N, bool GRAD, class Ex>
SY_HD void split_spectrum(Ex& ex, double* sm, int metric, const double* wsum_w, double* scratch, int64_t cap,
                          int64_t slot) {
  typedef LayoutJ<N> L;
  SY_STAGE_BEGIN(ex)
  if (g == 0) sm[L::FLAG] = 0.0;
  SY_STAGE_END(ex)
  w_from_pq<N>(ex, Scratch<N>::pb(scratch, cap, slot), Scratch<N>::qb(scratch, cap, slot), N, sm + L::GR, sm + L::GI);
  spectrum<N, L>(ex, sm, metric, wsum_w);
  if (GRAD) {
    park<N>(ex, sm + L::GR, Scratch<N>::gr(scratch, cap, slot));
    park<N>(ex, sm + L::GI, Scratch<N>::gi(scratch, cap, slot));
    SY_STAGE_BEGIN(ex)
    for (int k = g; k < N; k += L::G) Scratch<N>::coef(scratch, cap, slot)[k] = sm[L::COEF + k];
    SY_STAGE_END(ex)
  }
}
// part 3: fetch everything, backward
template <int N, class Ex>
SY_HD void split_backward(Ex& ex, double* sm, const double* p1, const double* p2, double* scratch, int64_t cap,
                          int64_t slot) {
  typedef Layout<N> L;
  fetch<N>(ex, Scratch<N>::li(scratch, cap, slot), sm + L::LI);
  fetch<N>(ex, Scratch<N>::pb(scratch, cap, slot), sm + L::PBUF);
  fetch<N>(ex, Scratch<N>::qb(scratch, cap, slot), sm + L::QBUF);
  fetch<N>(ex, Scratch<N>::gr(scratch, cap, slot), sm + L::GR);
  fetch<N>(ex, Scratch<N>::gi(scratch, cap, slot), sm + L::GI);
  SY_STAGE_BEGIN(ex)
  for (int k = g; k < N; k += L::G) sm[L::COEF + k] = Scratch<N>::coef(scratch, cap, slot)[k];
  SY_STAGE_END(ex)
  const GlobalIn<N> in{p1, p2};
  upper_backward<N>(ex, sm, in);
}

// fused: everything in one go (what the single-kernel path and the forward-only path use)
template <int N, bool GRAD, class Ex>
SY_HD void upper_pair(Ex& ex, double* sm, const double* p1, const double* p2, int metric, const double* wsum_w) {
  typedef Layout<N> L;
  const GlobalIn<N> in{p1, p2};
  upper_prologue<N>(ex, sm, in);
  w_from_pq<N>(ex, sm + L::PBUF, sm + L::QBUF, L::LD, sm + L::GR, sm + L::GI);
  spectrum<N, L>(ex, sm, metric, wsum_w);
  if (GRAD) upper_backward<N>(ex, sm, in);
}

// ---------------------------------------------------------------------------------------------
// Bounded domain, one pair, cooperative (bounded_domain.py:27-39): inverse Cayley transform of both
// operands by the same two-real-solves inverse, the upper half space pipeline on operands that live
// in shared memory, and the chain rule back through N_k = (I - z_k)^-1:
//   G_zk = conj(N_k) (2 G_Yk - 2i G_Xk) conj(N_k),  conj(N_k) = U_k - i V_k.
// Five buffers more than the upper half space layout (U1, V1, U2, V2 and a spare).
// Results: sm[DIST], sm[VS..], sm[FLAG]; GRAD: d dist / d z1 = (sm[G1R], sm[G1I]), d/dz2 = (sm[G2R], sm[G2I]).
template <int N>
struct BoundedLayout {
  typedef Layout<N> L;
  static constexpr int NN = L::NN, LD = L::LD, BUF = L::BUF, G = L::G;
  // memory: [U1 | V1 | U2 | V2 | SP] followed by a complete Layout<N> window (7 buffers + small state),
  // on which the shared upper-half pipeline runs unchanged
  static constexpr int U1 = 0, V1 = BUF, U2 = 2 * BUF, V2 = 3 * BUF, SP = 4 * BUF, UP = 5 * BUF;
  static constexpr int DIST = UP + L::DIST, VS = UP + L::VS, FLAG = UP + L::FLAG;
  static constexpr int G1R = UP + L::A3, G1I = SP, G2R = UP + L::A5, G2I = UP + L::LI;
  static constexpr int kRaw = UP + L::kRaw;
  static constexpr int kDoubles = (kRaw % 2 == 0) ? kRaw + 1 : kRaw;
};

template <int N, bool GRAD, class Ex>
SY_HD void bounded_pair(Ex& ex, double* sm, const double* p1, const double* p2, int metric, const double* wsum_w) {
  typedef BoundedLayout<N> BL;
  typedef Layout<N> L;
  constexpr int G = BL::G;
  constexpr int NN = BL::NN;
  constexpr int LD = BL::LD;
  double* up = sm + BL::UP;  // Layout<N> window: 7 buffers + small state
  double* u1 = sm + BL::U1, *v1 = sm + BL::V1, *u2 = sm + BL::U2, *v2 = sm + BL::V2, *sp = sm + BL::SP;
  double* a0 = up + L::A0, *a1 = up + L::A1, *a2 = up + L::A2, *a3 = up + L::A3, *a4 = up + L::A4, *a5 = up + L::A5;
  double* li = up + L::LI;
  double* rd = up + L::RD;
  double* flag = up + L::FLAG;

  SY_STAGE_BEGIN(ex)
  if (g == 0) *flag = 0.0;
  SY_STAGE_END(ex)
  for (int k = 0; k < 2; ++k) {
    const double* z = k == 0 ? p1 : p2;
    double* u = k == 0 ? u1 : u2;
    double* v = k == 0 ? v1 : v2;
    SY_STAGE_BEGIN(ex)
    for (int e = g; e < NN; e += G) {
      const int i = e / N, j = e - i * N, et = j * N + i, s = i * LD + j;
      a0[s] = ((i == j) ? 1.0 : 0.0) - 0.5 * (SY_LDG(z + e) + SY_LDG(z + et));  // E = I - Re z
      a1[s] = -0.5 * (SY_LDG(z + NN + e) + SY_LDG(z + NN + et));                  // F = -Im z
    }
    SY_STAGE_END(ex)
    inv_spd_real<N>(ex, a0, a1, u, v, a2, a3, a4, rd, flag);
  }
  const CayleyIn<N> in{u1, v1, u2, v2};
  upper_prologue<N>(ex, up, in);
  w_from_pq<N>(ex, up + L::PBUF, up + L::QBUF, L::LD, up + L::GR, up + L::GI);
  spectrum<N, L>(ex, up, metric, wsum_w);
  if (!GRAD) return;
  upper_backward<N>(ex, up, in);
  // upper results: G_X2 = a4, G_Y2 = a3, G_Y1 = a0, G_X1 = -a4.   Free: a1, a2, a5, li.
  // operand 2:  Gamma = 2 G_Y2 - 2i G_X2;  T = conj(N2) Gamma -> (a1, a2)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N], z0[N], z1[N], w0[N], w1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  zero2<N>(z0, z1);
  zero2<N>(w0, w1);
  mm_cols<N, false, false>(u2, a3, j0, j1, x0, x1);  // U Gy
  mm_cols<N, false, false>(v2, a4, j0, j1, y0, y1);  // V Gx
  mm_cols<N, false, false>(u2, a4, j0, j1, z0, z1);  // U Gx
  mm_cols<N, false, false>(v2, a3, j0, j1, w0, w1);  // V Gy
#pragma unroll
  for (int i = 0; i < N; ++i) {
    // (U - iV)(2Gy - 2iGx) = 2(U Gy - V Gx) - 2i(U Gx + V Gy)
    a1[i * LD + j0] = 2.0 * (x0[i] - y0[i]);
    a2[i * LD + j0] = -2.0 * (z0[i] + w0[i]);
    if (j1 < N) {
      a1[i * LD + j1] = 2.0 * (x1[i] - y1[i]);
      a2[i * LD + j1] = -2.0 * (z1[i] + w1[i]);
    }
  }
  SY_STAGE_END(ex)
  // G_z2 = T conj(N2) = (Tr + iTi)(U - iV) = (Tr U + Ti V) + i(Ti U - Tr V) -> (a5, li)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  mm_cols<N, false, false>(a1, u2, j0, j1, x0, x1);
  mm_cols<N, false, false>(a2, v2, j0, j1, x0, x1);
  mm_cols<N, false, false>(a2, u2, j0, j1, y0, y1);
  put_cols<N>(a5, j0, j1, x0, x1, 1.0, 0.0);
  double z0[N], z1[N];
  zero2<N>(z0, z1);
  mm_cols<N, false, false>(a1, v2, j0, j1, z0, z1);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    li[i * LD + j0] = y0[i] - z0[i];
    if (j1 < N) li[i * LD + j1] = y1[i] - z1[i];
  }
  SY_STAGE_END(ex)
  // operand 1:  Gamma = 2 G_Y1 - 2i G_X1 = 2 a0 + 2i a4;  T = conj(N1) Gamma -> (a1, a2)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N], z0[N], z1[N], w0[N], w1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  zero2<N>(z0, z1);
  zero2<N>(w0, w1);
  mm_cols<N, false, false>(u1, a0, j0, j1, x0, x1);  // U Gy
  mm_cols<N, false, false>(v1, a4, j0, j1, y0, y1);  // V Gx2
  mm_cols<N, false, false>(u1, a4, j0, j1, z0, z1);  // U Gx2
  mm_cols<N, false, false>(v1, a0, j0, j1, w0, w1);  // V Gy
#pragma unroll
  for (int i = 0; i < N; ++i) {
    // (U - iV)(2Gy + 2iGx2) = 2(U Gy + V Gx2) + 2i(U Gx2 - V Gy)
    a1[i * LD + j0] = 2.0 * (x0[i] + y0[i]);
    a2[i * LD + j0] = 2.0 * (z0[i] - w0[i]);
    if (j1 < N) {
      a1[i * LD + j1] = 2.0 * (x1[i] + y1[i]);
      a2[i * LD + j1] = 2.0 * (z1[i] - w1[i]);
    }
  }
  SY_STAGE_END(ex)
  // G_z1 = T conj(N1) -> (a3, sp)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double x0[N], x1[N], y0[N], y1[N], z0[N], z1[N];
  zero2<N>(x0, x1);
  zero2<N>(y0, y1);
  zero2<N>(z0, z1);
  mm_cols<N, false, false>(a1, u1, j0, j1, x0, x1);
  mm_cols<N, false, false>(a2, v1, j0, j1, x0, x1);
  mm_cols<N, false, false>(a2, u1, j0, j1, y0, y1);
  mm_cols<N, false, false>(a1, v1, j0, j1, z0, z1);
  put_cols<N>(a3, j0, j1, x0, x1, 1.0, 0.0);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    sp[i * LD + j0] = y0[i] - z0[i];
    if (j1 < N) sp[i * LD + j1] = y1[i] - z1[i];
  }
  SY_STAGE_END(ex)
}

// ---------------------------------------------------------------------------------------------
// SPD (affine-invariant) distance, one pair, cooperative: dist^2 = sum log^2 lambda_i(X^-1 Y),
// lambda = sigma^2 of Gm = Lx^-1 Ly (see spd_pair in pair_math_impl.inc).  p1 / p2 point at the
// (n, n) rows.  Results: sm[DIST], sm[VS..], sm[FLAG]; GRAD: d dist / dX = sm[SPD_GX], d/dY = sm[SPD_GY].
template <int N>
struct SpdLayout : Layout<N> {
  typedef Layout<N> B;
  static constexpr int XI = B::LI, LY = B::A0, LYI = B::A1, GM = B::A2, GJ = B::A3, T0 = B::A4, T1 = B::A5;
  static constexpr int SPD_GX = T0, SPD_GY = T1;
};

template <int N, bool GRAD, class Ex>
SY_HD void spd_pair(Ex& ex, double* sm, const double* p1, const double* p2) {
  typedef SpdLayout<N> L;
  constexpr int G = L::G;
  constexpr int NN = L::NN;
  constexpr int LD = L::LD;
  double* xi = sm + L::XI;
  double* ly = sm + L::LY;
  double* lyi = sm + L::LYI;
  double* gm = sm + L::GM;
  double* gj = sm + L::GJ;
  double* t0 = sm + L::T0;
  double* t1 = sm + L::T1;
  double* rd = sm + L::RD;
  double* flag = sm + L::FLAG;

  SY_STAGE_BEGIN(ex)
  if (g == 0) *flag = 0.0;
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N, et = j * N + i, s = i * LD + j;
    t0[s] = 0.5 * (SY_LDG(p1 + e) + SY_LDG(p1 + et));
    t1[s] = 0.5 * (SY_LDG(p2 + e) + SY_LDG(p2 + et));
  }
  SY_STAGE_END(ex)
  chol_inv<N>(ex, t0, gj, xi, rd, flag);         // xi = chol(X)^-1
  chol_inv<N>(ex, t1, ly, lyi, rd, flag, true);  // ly = chol(Y), lyi = its inverse
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(xi, ly, gm, g, 1.0, 0.0);  // Gm = Lx^-1 Ly (lower triangular)
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int c = e / N, r = e - c * N;
    gj[c * LD + r] = gm[r * LD + c];  // column-major copy for the Jacobi
  }
  SY_STAGE_END(ex)
  const int sweeps = ex.template jacobi<N, true>(gj, gj, sm + L::WORST);
  SY_STAGE_BEGIN(ex)
  for (int k = g; k < N; k += G) {
    double s0 = 0.0;
    for (int i = 0; i < N; ++i) s0 += gj[k * LD + i] * gj[k * LD + i];
    const double sig = sqrt(s0);
    sm[L::SIG + k] = sig;
    sm[L::VK + k] = 2.0 * log(sig);
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double ss = 0.0;
  for (int k = 0; k < N; ++k) ss += sm[L::VK + k] * sm[L::VK + k];
  const double dist = sqrt(ss);
  const double inv = dist > 0.0 ? 1.0 / dist : 0.0;
  for (int k = g; k < N; k += G) {
    const double d = sm[L::SIG + k];
    int r = 0;
    for (int j = 0; j < N; ++j) {
      const double o = sm[L::SIG + j];
      r += (o < d || (o == d && j < k)) ? 1 : 0;
    }
    sm[L::VS + r] = sm[L::VK + k];
    const double s2 = d * d;
    sm[L::COEF + k] = sm[L::VK + k] * inv * 2.0 / (s2 * s2);  // (dL/dsigma_k) / sigma_k^3
  }
  if (g == 0) {
    unsigned st = (*flag != 0.0) ? kStatusNotPD : 0u;
    if (sweeps >= kMaxSweeps) st |= kStatusNoConverge;
    if (!(dist == dist) || dist > 1e300) st |= kStatusNonFinite;
    sm[L::DIST] = dist;
    *flag = (double)st;
  }
  SY_STAGE_END(ex)
  if (!GRAD) return;

  // F = sum_k coef_k g_k g_k^T -> t0
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  const bool two = j1 < N;
  const int jj1 = two ? j1 : j0;
  double f0[N], f1[N];
  zero2<N>(f0, f1);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double c = sm[L::COEF + k];
    const double b0 = c * gj[k * LD + j0];
    const double b1 = two ? c * gj[k * LD + jj1] : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double a = gj[k * LD + i];
      f0[i] += a * b0;
      f1[i] += a * b1;
    }
  }
  put_cols<N>(t0, j0, j1, f0, f1, 1.0, 0.0);
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(t0, gm, t1, g, 1.0, 0.0);  // t1 = G_Gm = F Gm
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(xi, t1, gj, g, 1.0, 0.0);   // gj = Lxi^T G_Gm   (lower part = G_Ly)
  mm_stage<N, false, true>(t1, ly, t0, g, 1.0, 0.0);   // t0 = G_Gm Ly^T    (lower part = G_Lxi)
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(ly, gj, t1, g, 1.0, 0.0);   // t1 = Ly^T G_Ly    (lower part used)
  mm_stage<N, false, true>(t0, xi, gm, g, 1.0, 0.0);   // gm = G_Lxi Lxi^T  (lower part used)
  SY_STAGE_END(ex)
  // Phi = 0.5 tril(.) mirrored (in place: the upper triangle is written only by the owner of (i, j))
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N;
    if (i >= j) {
      const double a = 0.5 * t1[i * LD + j], b = 0.5 * gm[i * LD + j];
      t1[i * LD + j] = a;
      t1[j * LD + i] = a;
      gm[i * LD + j] = b;
      gm[j * LD + i] = b;
    }
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, false, false>(t1, lyi, gj, g, 1.0, 0.0);  // gj = Phi Lyi
  mm_stage<N, false, false>(gm, xi, t0, g, 1.0, 0.0);   // t0 = K_s Lxi
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  mm_stage<N, true, false>(lyi, gj, t1, g, 1.0, 0.0);   // t1 = G_Y = Lyi^T Phi Lyi
  mm_stage<N, true, false>(xi, t0, gm, g, -1.0, 0.0);   // gm = G_X = -Lxi^T K_s Lxi
  SY_STAGE_END(ex)
  // move G_X where the layout promises it (t0)
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e - i * N;
    t0[i * LD + j] = gm[i * LD + j];
  }
  SY_STAGE_END(ex)
}

}  // namespace coop
}  // namespace sympa
