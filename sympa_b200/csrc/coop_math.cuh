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

template <int N>
struct Layout {
  static constexpr int NN = N * N;
  static constexpr int NP = (N % 2 == 0) ? N : N + 1;  // padded column count for the tournament
  static constexpr int G = NP / 2;                      // lanes per pair
  // buffer offsets (doubles) inside one pair's region
  static constexpr int LI = 0, P = NN, Q = 2 * NN, GR = 3 * NN, GI = 4 * NN, T0 = 5 * NN, T1 = 6 * NN, T2 = 7 * NN;
  static constexpr int SIG = 8 * NN, COEF = SIG + N, VS = COEF + N, RD = VS + N, WORST = RD + N;  // WORST: 2*G
  static constexpr int DIST = WORST + 2 * G, FLAG = DIST + 1;
  static constexpr int kRaw = FLAG + 1;
  static constexpr int kDoubles = (kRaw % 2 == 0) ? kRaw + 1 : kRaw;  // odd stride: spreads banks across pair slots
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
};

// ---------------------------------------------------------------------------------------------
// dense helper: accumulate the two owned columns j0 = g, j1 = g + G (j1 may be >= N: ignored)
//   acc0[i] += sum_k A(i,k) * B(k, j0)      A(i,k) = TA ? a[k*N + i] : a[i*N + k]
//                                           B(k,j) = TB ? b[j*N + k] : b[k*N + j]
template <int N, bool TA, bool TB>
SY_HD void mm_cols(const double* a, const double* b, int j0, int j1, double* acc0, double* acc1) {
  const bool two = j1 < N;
  const int jj1 = two ? j1 : j0;
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    const double b0 = TB ? b[j0 * N + k] : b[k * N + j0];
    double b1 = TB ? b[jj1 * N + k] : b[k * N + jj1];
    b1 = two ? b1 : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double av = TA ? a[k * N + i] : a[i * N + k];
      acc0[i] += av * b0;
      acc1[i] += av * b1;
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

// write owned columns: c[i*N + j] = s * acc[i] (+ diag on the diagonal)
template <int N>
SY_HD void put_cols(double* c, int j0, int j1, const double* acc0, const double* acc1, double s, double diag) {
#pragma unroll
  for (int i = 0; i < N; ++i) c[i * N + j0] = s * acc0[i] + ((i == j0) ? diag : 0.0);
  if (j1 < N) {
#pragma unroll
    for (int i = 0; i < N; ++i) c[i * N + j1] = s * acc1[i] + ((i == j1) ? diag : 0.0);
  }
}

// Inverse Cholesky factor, cooperative: a (full symmetric, lower part used) -> li (full, upper part
// zeroed); l = scratch (full), rd = n doubles.  N + 1 stages.
template <int N, class Ex>
SY_HD void chol_inv(Ex& ex, const double* a, double* l, double* li, double* rd, double* flag) {
  constexpr int G = Layout<N>::G;
#pragma unroll 1
  for (int j = 0; j < N; ++j) {
    SY_STAGE_BEGIN(ex)
    double d = a[j * N + j];
    for (int k = 0; k < j; ++k) d -= l[j * N + k] * l[j * N + k];
    const double r = sy_rsqrt(d);
    for (int i = j + g; i < N; i += G) {
      if (i == j) {
        l[j * N + j] = d * r;
        rd[j] = r;
        if (!(d > 0.0)) *flag = 1.0;
      } else {
        double s = a[i * N + j];
        for (int k = 0; k < j; ++k) s -= l[i * N + k] * l[j * N + k];
        l[i * N + j] = s * r;
      }
    }
    SY_STAGE_END(ex)
  }
  SY_STAGE_BEGIN(ex)
  for (int c = g; c < N; c += G) {
    for (int i = 0; i < c; ++i) li[i * N + c] = 0.0;
    li[c * N + c] = rd[c];
    for (int i = c + 1; i < N; ++i) {
      double s = 0.0;
      for (int k = c; k < i; ++k) s += l[i * N + k] * li[k * N + c];
      li[i * N + c] = -s * rd[i];
    }
  }
  SY_STAGE_END(ex)
}

// (E + iF)^-1 = U + iV, E SPD (buffer e), F symmetric (buffer f).  Scratch: s0, s1, s2 (full
// buffers, all distinct from e, f, u, v and from each other; u may not alias anything).
//   lei -> s0 ; T = lei F -> s1 ; S = E + T^T T -> s2 ; lsi -> e (E dead) ; U = lsi^T lsi -> u ;
//   TU -> s2 ; V = -lei^T TU -> v
template <int N, class Ex>
SY_HD void inv_spd_real(Ex& ex, double* e, const double* f, double* u, double* v, double* s0, double* s1, double* s2,
                        double* rd, double* flag) {
  constexpr int G = Layout<N>::G;
  chol_inv<N>(ex, e, s1, s0, rd, flag);  // s1 = scratch L, s0 = lei
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(s0, f, g, g + G, a0, a1);
  put_cols<N>(s1, g, g + G, a0, a1, 1.0, 0.0);  // T
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(s1, s1, g, g + G, a0, a1);  // T^T T
#pragma unroll
  for (int i = 0; i < N; ++i) {
    s2[i * N + g] = a0[i] + e[i * N + g];
    if (g + G < N) s2[i * N + g + G] = a1[i] + e[i * N + g + G];
  }
  SY_STAGE_END(ex)
  chol_inv<N>(ex, s2, u, e, rd, flag);  // u = scratch L, e = lsi
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(e, e, g, g + G, a0, a1);  // U = lsi^T lsi
  put_cols<N>(u, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(s1, u, g, g + G, a0, a1);  // TU
  put_cols<N>(s2, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(s0, s2, g, g + G, a0, a1);  // lei^T TU
  put_cols<N>(v, g, g + G, a0, a1, -1.0, 0.0);
  SY_STAGE_END(ex)
}

// tournament pairing: round r in [0, NP-1), slot k in [0, NP/2)
template <int N>
SY_HD void tournament(int r, int k, int* p, int* q) {
  constexpr int NP = Layout<N>::NP;
  int a, b;
  if (k == 0) {
    a = NP - 1;
    b = r;
  } else {
    a = (r + k) % (NP - 1);
    b = (r - k + (NP - 1)) % (NP - 1);
  }
  *p = a < b ? a : b;
  *q = a < b ? b : a;
}

// One-sided Jacobi on the column-major complex matrix (gr, gi) in shared memory (no V).
template <int N, class Ex>
SY_HD int jacobi(Ex& ex, double* gr, double* gi, double* worst) {
  constexpr int NP = Layout<N>::NP;
  constexpr int G = Layout<N>::G;
  int sweep = 0;
#pragma unroll 1
  for (; sweep < kMaxSweeps; ++sweep) {
    double* wbuf = worst + (sweep & 1) * G;
#pragma unroll 1
    for (int r = 0; r < NP - 1; ++r) {
      SY_STAGE_BEGIN(ex)
      int p, q;
      tournament<N>(r, g, &p, &q);
      double ratio = 0.0;
      if (q < N) {  // q == N is the padding column of an odd n
        double pr[N], pi[N], qr[N], qi[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          pr[i] = gr[p * N + i];
          pi[i] = gi[p * N + i];
          qr[i] = gr[q * N + i];
          qi[i] = gi[q * N + i];
        }
        double al = 0.0, be = 0.0, cr = 0.0, ci = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          al += pr[i] * pr[i] + pi[i] * pi[i];
          be += qr[i] * qr[i] + qi[i] * qi[i];
          cr += pr[i] * qr[i] + pi[i] * qi[i];
          ci += pr[i] * qi[i] - pi[i] * qr[i];
        }
        const double g2 = cr * cr + ci * ci;
        const double ab = al * be;
        const bool rot = g2 > reg::kSkipRatio2 * ab;
        if (rot) {
          ratio = g2 / ab;
          const double dl = 0.5 * (be - al);
          const double h = dl * dl + g2;
          const double den = fabs(dl) + h * sy_rsqrt(h);
          const double rden = sy_rsqrt(den);
          double inv = rden * rden;
          inv = dl < 0.0 ? -inv : inv;
          const double tr_ = cr * inv, ti_ = ci * inv;
          const double c = sy_rsqrt(1.0 + tr_ * tr_ + ti_ * ti_);
          const double sr = c * tr_, si = c * ti_;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            gr[p * N + i] = c * pr[i] - (sr * qr[i] + si * qi[i]);
            gi[p * N + i] = c * pi[i] - (sr * qi[i] - si * qr[i]);
            gr[q * N + i] = c * qr[i] + (sr * pr[i] - si * pi[i]);
            gi[q * N + i] = c * qi[i] + (sr * pi[i] + si * pr[i]);
          }
        }
      }
      wbuf[g] = (r == 0) ? ratio : (ratio > wbuf[g] ? ratio : wbuf[g]);
      SY_STAGE_END(ex)
    }
    double w = 0.0;
    for (int k = 0; k < G; ++k) w = wbuf[k] > w ? wbuf[k] : w;
    const bool more = !(w < reg::kStopRatio2);
    if (!ex.any(more)) {
      ++sweep;
      break;
    }
  }
  return sweep;
}

// ---------------------------------------------------------------------------------------------
// Upper half space, one pair, cooperative.  p1 / p2 point at the (2, n, n) rows in global memory.
// Results are left in the pair's shared-memory region:
//   sm[DIST], sm[VS .. VS+N) (ascending vvd), sm[FLAG] (status bits as a double)
//   GRAD:  d dist / d X2 = sm[Q],  d/dY2 = sm[T2],  d/dY1 = sm[P],  d/dX1 = -sm[Q]   (full n x n)
template <int N, bool GRAD, class Ex>
SY_HD void upper_pair(Ex& ex, double* sm, const double* p1, const double* p2, int metric, const double* wsum_w) {
  typedef Layout<N> L;
  constexpr int G = L::G;
  constexpr int NN = L::NN;
  double* li = sm + L::LI;
  double* P = sm + L::P;
  double* Q = sm + L::Q;
  double* gr = sm + L::GR;
  double* gi = sm + L::GI;
  double* t0 = sm + L::T0;
  double* t1 = sm + L::T1;
  double* t2 = sm + L::T2;
  double* rd = sm + L::RD;
  double* flag = sm + L::FLAG;

  // ---- load: t0 = D = X2 - X1, t1 = Y1, t2 = Y2 (symmetrised)
  SY_STAGE_BEGIN(ex)
  if (g == 0) *flag = 0.0;
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e % N, et = j * N + i;
    t0[e] = 0.5 * ((SY_LDG(p2 + e) + SY_LDG(p2 + et)) - (SY_LDG(p1 + e) + SY_LDG(p1 + et)));
    t1[e] = 0.5 * (SY_LDG(p1 + NN + e) + SY_LDG(p1 + NN + et));
    t2[e] = 0.5 * (SY_LDG(p2 + NN + e) + SY_LDG(p2 + NN + et));
  }
  SY_STAGE_END(ex)

  chol_inv<N>(ex, t1, Q, li, rd, flag);  // li = chol(Y1)^-1

  // ---- t1 = li D, gr = li Y2
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(li, t0, g, g + G, a0, a1);
  put_cols<N>(t1, g, g + G, a0, a1, 1.0, 0.0);
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(li, t2, g, g + G, a0, a1);
  put_cols<N>(gr, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  // ---- t0 = -A = -(li D) li^T ; t2 = C = (li Y2) li^T + I
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, true>(t1, li, g, g + G, a0, a1);
  put_cols<N>(t0, g, g + G, a0, a1, -1.0, 0.0);
  zero2<N>(a0, a1);
  mm_cols<N, false, true>(gr, li, g, g + G, a0, a1);
  put_cols<N>(t2, g, g + G, a0, a1, 1.0, 1.0);
  SY_STAGE_END(ex)

  // ---- (C - iA)^-1 = Q + iP   ->   (A + iC)^-1 = P - iQ
  inv_spd_real<N>(ex, t2, t0, Q, P, gi, t1, gr, rd, flag);

  // ---- W = (I - 2Q) - 2iP, column-major for the Jacobi
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int c = e / N, r = e % N;
    gr[e] = ((r == c) ? 1.0 : 0.0) - 2.0 * Q[r * N + c];
    gi[e] = -2.0 * P[r * N + c];
  }
  SY_STAGE_END(ex)

  const int sweeps = jacobi<N>(ex, gr, gi, sm + L::WORST);

  // ---- singular values, ranking, vector-valued distance, metric (one lane; O(n^2) scalar work)
  SY_STAGE_BEGIN(ex)
  for (int k = g; k < N; k += G) {
    double a = 0.0;
    for (int i = 0; i < N; ++i) a += gr[k * N + i] * gr[k * N + i] + gi[k * N + i] * gi[k * N + i];
    sm[L::SIG + k] = sqrt(a);
  }
  SY_STAGE_END(ex)
  SY_STAGE_BEGIN(ex)
  if (g == 0) {
    double sig[N], v[N], gsig[N];
    int rank[N];
    unsigned st = (*flag != 0.0) ? kStatusNotPD : 0u;
    if (sweeps >= kMaxSweeps) st |= kStatusNoConverge;
    for (int k = 0; k < N; ++k) sig[k] = sm[L::SIG + k];
    const double dist = loc::metric_reduce<N>(sig, metric, wsum_w, rank, v, gsig, &st);
    if (!(dist == dist) || dist > 1e300) st |= kStatusNonFinite;
    for (int k = 0; k < N; ++k) {
      sm[L::VS + rank[k]] = v[k];
      // dL/dsigma_k / sigma_k^3   (v_k = W^H g_k / sigma_k^2, u_k = g_k / sigma_k)
      sm[L::COEF + k] = sig[k] > 1e-150 ? gsig[k] / (sig[k] * sig[k] * sig[k]) : 0.0;
    }
    sm[L::DIST] = dist;
    *flag = (double)st;
  }
  SY_STAGE_END(ex)
  if (!GRAD) return;

  // ---- F = sum_k coef_k g_k g_k^H  (Hermitian):  t0 = Re F, t1 = Im F
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
    const double b0r = c * gr[k * N + j0], b0i = c * gi[k * N + j0];
    const double b1r = two ? c * gr[k * N + jj1] : 0.0, b1i = two ? c * gi[k * N + jj1] : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double ar = gr[k * N + i], ai = gi[k * N + i];
      // g_k[i] * conj(g_k[j])
      fr0[i] += ar * b0r + ai * b0i;
      fi0[i] += ai * b0r - ar * b0i;
      fr1[i] += ar * b1r + ai * b1i;
      fi1[i] += ai * b1r - ar * b1i;
    }
  }
  put_cols<N>(t0, j0, j1, fr0, fr1, 1.0, 0.0);
  put_cols<N>(t1, j0, j1, fi0, fi1, 1.0, 0.0);
  SY_STAGE_END(ex)

  // ---- GW = F W,  W = (I - 2Q) - 2iP   ->  gr = Re GW, gi = Im GW (row-major from here on)
  //      F W = F - 2 F (Q + iP):   Re = Fr - 2 (Fr Q - Fi P),  Im = Fi - 2 (Fr P + Fi Q)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double a0[N], a1[N], b0[N], b1[N];
  zero2<N>(a0, a1);
  zero2<N>(b0, b1);
  mm_cols<N, false, false>(t0, Q, j0, j1, a0, a1);  // Fr Q
  mm_cols<N, false, false>(t0, P, j0, j1, b0, b1);  // Fr P
  double c0[N], c1[N], d0[N], d1[N];
  zero2<N>(c0, c1);
  zero2<N>(d0, d1);
  mm_cols<N, false, false>(t1, P, j0, j1, c0, c1);  // Fi P
  mm_cols<N, false, false>(t1, Q, j0, j1, d0, d1);  // Fi Q
#pragma unroll
  for (int i = 0; i < N; ++i) {
    gr[i * N + j0] = t0[i * N + j0] - 2.0 * (a0[i] - c0[i]);
    gi[i * N + j0] = t1[i * N + j0] - 2.0 * (b0[i] + d0[i]);
    if (j1 < N) {
      gr[i * N + j1] = t0[i * N + j1] - 2.0 * (a1[i] - c1[i]);
      gi[i * N + j1] = t1[i * N + j1] - 2.0 * (b1[i] + d1[i]);
    }
  }
  SY_STAGE_END(ex)

  // ---- G_N = 2i sym(GW) = i (GW + GW^T):  t0 = Re = -(GWi + GWi^T),  t1 = Im = GWr + GWr^T
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e % N, et = j * N + i;
    t0[e] = -(gi[e] + gi[et]);
    t1[e] = gr[e] + gr[et];
  }
  SY_STAGE_END(ex)

  // ---- Tm = conj(N) G_N, conj(N) = P + iQ:  gr = P GNr - Q GNi,  gi = P GNi + Q GNr
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double a0[N], a1[N], b0[N], b1[N], c0[N], c1[N], d0[N], d1[N];
  zero2<N>(a0, a1);
  zero2<N>(b0, b1);
  zero2<N>(c0, c1);
  zero2<N>(d0, d1);
  mm_cols<N, false, false>(P, t0, j0, j1, a0, a1);
  mm_cols<N, false, false>(Q, t1, j0, j1, b0, b1);
  mm_cols<N, false, false>(P, t1, j0, j1, c0, c1);
  mm_cols<N, false, false>(Q, t0, j0, j1, d0, d1);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    gr[i * N + j0] = a0[i] - b0[i];
    gi[i * N + j0] = c0[i] + d0[i];
    if (j1 < N) {
      gr[i * N + j1] = a1[i] - b1[i];
      gi[i * N + j1] = c1[i] + d1[i];
    }
  }
  SY_STAGE_END(ex)

  // ---- G_M = -Tm conj(N):  t0 = G_A = -(Tmr P - Tmi Q),  t1 = G_B = -(Tmr Q + Tmi P)
  SY_STAGE_BEGIN(ex)
  const int j0 = g, j1 = g + G;
  double a0[N], a1[N], b0[N], b1[N], c0[N], c1[N], d0[N], d1[N];
  zero2<N>(a0, a1);
  zero2<N>(b0, b1);
  zero2<N>(c0, c1);
  zero2<N>(d0, d1);
  mm_cols<N, false, false>(gr, P, j0, j1, a0, a1);
  mm_cols<N, false, false>(gi, Q, j0, j1, b0, b1);
  mm_cols<N, false, false>(gr, Q, j0, j1, c0, c1);
  mm_cols<N, false, false>(gi, P, j0, j1, d0, d1);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    t0[i * N + j0] = -(a0[i] - b0[i]);
    t1[i * N + j0] = -(c0[i] + d0[i]);
    if (j1 < N) {
      t0[i * N + j1] = -(a1[i] - b1[i]);
      t1[i * N + j1] = -(c1[i] + d1[i]);
    }
  }
  SY_STAGE_END(ex)

  // ---- symmetrise G_A, G_B in place is not possible inside one stage: use them through sym access
  //      below (0.5 (g[i][j] + g[j][i])) only where symmetry matters (it does not: every use is a
  //      congruence or a product whose result is symmetrised at the end).

  // ---- reload D -> Q, Y2 -> P (P, Q are dead)
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e % N, et = j * N + i;
    Q[e] = 0.5 * ((SY_LDG(p2 + e) + SY_LDG(p2 + et)) - (SY_LDG(p1 + e) + SY_LDG(p1 + et)));
    P[e] = 0.5 * (SY_LDG(p2 + NN + e) + SY_LDG(p2 + NN + et));
  }
  SY_STAGE_END(ex)
  // ---- gr = T1 = li D, gi = T2 = li Y2
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(li, Q, g, g + G, a0, a1);
  put_cols<N>(gr, g, g + G, a0, a1, 1.0, 0.0);
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(li, P, g, g + G, a0, a1);
  put_cols<N>(gi, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  // ---- t2 = G_Li = 2 (G_A T1 + G_B T2)   (only its lower triangle is used)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(t0, gr, g, g + G, a0, a1);
  mm_cols<N, false, false>(t1, gi, g, g + G, a0, a1);
  put_cols<N>(t2, g, g + G, a0, a1, 2.0, 0.0);
  SY_STAGE_END(ex)
  // ---- Q = K_s: 0.5 * tril(G_Li li^T) mirrored to a full symmetric matrix
  SY_STAGE_BEGIN(ex)
  for (int e = g; e < NN; e += G) {
    const int i = e / N, j = e % N;
    if (i >= j) {
      double a = 0.0;
      for (int k = 0; k <= j; ++k) a += t2[i * N + k] * li[j * N + k];
      Q[i * N + j] = 0.5 * a;
      Q[j * N + i] = 0.5 * a;
    }
  }
  SY_STAGE_END(ex)
  // ---- gr = K_s li
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(Q, li, g, g + G, a0, a1);
  put_cols<N>(gr, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  // ---- P = G_Y1 = -li^T (K_s li)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(li, gr, g, g + G, a0, a1);
  put_cols<N>(P, g, g + G, a0, a1, -1.0, 0.0);
  SY_STAGE_END(ex)
  // ---- gr = G_A li, gi = G_B li
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(t0, li, g, g + G, a0, a1);
  put_cols<N>(gr, g, g + G, a0, a1, 1.0, 0.0);
  zero2<N>(a0, a1);
  mm_cols<N, false, false>(t1, li, g, g + G, a0, a1);
  put_cols<N>(gi, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
  // ---- Q = G_X2 = li^T (G_A li), t2 = G_Y2 = li^T (G_B li)
  SY_STAGE_BEGIN(ex)
  double a0[N], a1[N];
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(li, gr, g, g + G, a0, a1);
  put_cols<N>(Q, g, g + G, a0, a1, 1.0, 0.0);
  zero2<N>(a0, a1);
  mm_cols<N, true, false>(li, gi, g, g + G, a0, a1);
  put_cols<N>(t2, g, g + G, a0, a1, 1.0, 0.0);
  SY_STAGE_END(ex)
}

}  // namespace coop
}  // namespace sympa
