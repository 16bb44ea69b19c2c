// TEST INFRASTRUCTURE: compiles sympa_b200/csrc/pair_math.cuh for the HOST so that the exact
// per-pair templates the CUDA kernels instantiate can be checked against the golden vectors on a
// machine without a GPU (tests/test_hostcheck.py).  Never loaded by the product package.
//   variant 0 = namespace sympa::reg (unrolled, what the kernels use for n <= 4)
//   variant 1 = namespace sympa::loc (rolled loops, per-thread fallback)
//   variant 2 = namespace sympa::coop (warp-cooperative shared-memory kernels, lanes emulated sequentially)
//   variant 3 = the same, split in three parts with the state parked in (poisoned) global scratch
#include "../../sympa_b200/csrc/pair_math.cuh"
#include "../../sympa_b200/csrc/coop_math.cuh"
#include <vector>

using namespace sympa;

#define HC_RUN(NS)                                                                                              \
  template <int N>                                                                                              \
  static void run_##NS(int kind, int metric, int64_t b, const double* z1, const double* z2, const double* w,    \
                       int grad, double* dist, double* vvd, double* g1, double* g2, unsigned* status) {        \
    constexpr int T = Cfg<N>::kTri;                                                                             \
    const int per = (kind == kSpd ? 1 : 2) * N * N;                                                             \
    for (int64_t p = 0; p < b; ++p) {                                                                           \
      const double* a = z1 + p * per;                                                                           \
      const double* c = z2 + p * per;                                                                           \
      double vs[N];                                                                                             \
      unsigned st = 0;                                                                                          \
      double d;                                                                                                 \
      double o1r[T], o1i[T], o2r[T], o2i[T];                                                                    \
      if (kind == kSpd) {                                                                                       \
        double x[T], y[T];                                                                                      \
        NS::pack_sym<N>(a, x);                                                                                  \
        NS::pack_sym<N>(c, y);                                                                                  \
        d = grad ? NS::spd_pair<N, true>(x, y, vs, o1r, o2r, &st) : NS::spd_pair<N, false>(x, y, vs, o1r, o2r, &st); \
      } else {                                                                                                  \
        double x1[T], y1[T], x2[T], y2[T];                                                                      \
        NS::pack_sym<N>(a, x1);                                                                                 \
        NS::pack_sym<N>(a + N * N, y1);                                                                         \
        NS::pack_sym<N>(c, x2);                                                                                 \
        NS::pack_sym<N>(c + N * N, y2);                                                                         \
        if (kind == kUpper)                                                                                     \
          d = grad ? NS::upper_pair<N, true>(x1, y1, x2, y2, metric, w, vs, o1r, o1i, o2r, o2i, &st)            \
                   : NS::upper_pair<N, false>(x1, y1, x2, y2, metric, w, vs, o1r, o1i, o2r, o2i, &st);          \
        else                                                                                                    \
          d = grad ? NS::bounded_pair<N, true>(x1, y1, x2, y2, metric, w, vs, o1r, o1i, o2r, o2i, &st)          \
                   : NS::bounded_pair<N, false>(x1, y1, x2, y2, metric, w, vs, o1r, o1i, o2r, o2i, &st);        \
      }                                                                                                         \
      dist[p] = d;                                                                                              \
      for (int k = 0; k < N; ++k) vvd[p * N + k] = vs[k];                                                       \
      *status |= st;                                                                                            \
      if (grad) {                                                                                               \
        for (int i = 0; i < N; ++i)                                                                             \
          for (int j = 0; j < N; ++j) {                                                                         \
            g1[p * per + i * N + j] = o1r[tri(i, j)];                                                           \
            g2[p * per + i * N + j] = o2r[tri(i, j)];                                                           \
            if (kind != kSpd) {                                                                                 \
              g1[p * per + N * N + i * N + j] = o1i[tri(i, j)];                                                 \
              g2[p * per + N * N + i * N + j] = o2i[tri(i, j)];                                                 \
            }                                                                                                   \
          }                                                                                                     \
      }                                                                                                         \
    }                                                                                                           \
  }

HC_RUN(reg)
HC_RUN(loc)

template <int N, bool SPLIT>
static void run_coop(int kind, int metric, int64_t b, const double* z1, const double* z2, const double* w, int grad,
                     double* dist, double* vvd, double* g1, double* g2, unsigned* status) {
  typedef coop::Layout<N> L;
  typedef coop::LayoutJ<N> LJ;
  if (kind == kSpd) {
    typedef coop::SpdLayout<N> S;
    std::vector<double> sm(S::kDoubles);
    coop::HostExec ex{S::G};
    const int per = N * N;
    for (int64_t p = 0; p < b; ++p) {
      for (auto& x : sm) x = -7.0e300;
      if (grad)
        coop::spd_pair<N, true>(ex, sm.data(), z1 + p * per, z2 + p * per);
      else
        coop::spd_pair<N, false>(ex, sm.data(), z1 + p * per, z2 + p * per);
      dist[p] = sm[S::DIST];
      for (int k = 0; k < N; ++k) vvd[p * N + k] = sm[S::VS + k];
      *status |= (unsigned)sm[S::FLAG];
      if (grad)
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) {
            const int s = i * S::LD + j, t = j * S::LD + i;
            g1[p * per + i * N + j] = 0.5 * (sm[S::SPD_GX + s] + sm[S::SPD_GX + t]);
            g2[p * per + i * N + j] = 0.5 * (sm[S::SPD_GY + s] + sm[S::SPD_GY + t]);
          }
    }
    return;
  }
  if (kind == kBounded) {
    typedef coop::BoundedLayout<N> S;
    std::vector<double> sm(S::kDoubles);
    coop::HostExec ex{S::G};
    const int per = 2 * N * N;
    for (int64_t p = 0; p < b; ++p) {
      for (auto& x : sm) x = -7.0e300;
      if (grad)
        coop::bounded_pair<N, true>(ex, sm.data(), z1 + p * per, z2 + p * per, metric, w);
      else
        coop::bounded_pair<N, false>(ex, sm.data(), z1 + p * per, z2 + p * per, metric, w);
      dist[p] = sm[S::DIST];
      for (int k = 0; k < N; ++k) vvd[p * N + k] = sm[S::VS + k];
      *status |= (unsigned)sm[S::FLAG];
      if (grad)
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) {
            const int s = i * S::LD + j, t = j * S::LD + i, e = i * N + j;
            g1[p * per + e] = 0.5 * (sm[S::G1R + s] + sm[S::G1R + t]);
            g1[p * per + N * N + e] = 0.5 * (sm[S::G1I + s] + sm[S::G1I + t]);
            g2[p * per + e] = 0.5 * (sm[S::G2R + s] + sm[S::G2R + t]);
            g2[p * per + N * N + e] = 0.5 * (sm[S::G2I + s] + sm[S::G2I + t]);
          }
    }
    return;
  }
  const int per = 2 * N * N;
  std::vector<double> sm(L::kDoubles), smj(LJ::kDoubles);
  const int64_t cap = 3;  // scratch capacity in pairs; the slot rotates so that the indexing is exercised
  std::vector<double> scratch(coop::Scratch<N>::kPerPair * cap);
  coop::HostExec ex{L::G};
  for (int64_t p = 0; p < b; ++p) {
    for (auto& x : sm) x = -7.0e300;  // poison: any read of an unwritten slot shows up
    for (auto& x : smj) x = -7.0e300;
    if (!SPLIT) {
      if (grad)
        coop::upper_pair<N, true>(ex, sm.data(), z1 + p * per, z2 + p * per, metric, w);
      else
        coop::upper_pair<N, false>(ex, sm.data(), z1 + p * per, z2 + p * per, metric, w);
      dist[p] = sm[L::DIST];
      for (int k = 0; k < N; ++k) vvd[p * N + k] = sm[L::VS + k];
      *status |= (unsigned)sm[L::FLAG];
    } else {
      for (auto& x : scratch) x = -7.0e300;
      const int64_t slot = p % cap;
      coop::split_prologue<N>(ex, sm.data(), z1 + p * per, z2 + p * per, scratch.data(), cap, slot);
      if (sm[L::FLAG] != 0.0) *status |= kStatusNotPD;
      for (auto& x : sm) x = -7.0e300;
      if (grad)
        coop::split_spectrum<N, true>(ex, smj.data(), metric, w, scratch.data(), cap, slot);
      else
        coop::split_spectrum<N, false>(ex, smj.data(), metric, w, scratch.data(), cap, slot);
      dist[p] = smj[LJ::DIST];
      for (int k = 0; k < N; ++k) vvd[p * N + k] = smj[LJ::VS + k];
      *status |= (unsigned)smj[LJ::FLAG];
      if (grad) coop::split_backward<N>(ex, sm.data(), z1 + p * per, z2 + p * per, scratch.data(), cap, slot);
    }
    if (grad) {
      const double* gx2 = &sm[L::GX2];
      const double* gy2 = &sm[L::GY2];
      const double* gy1 = &sm[L::GY1];
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          const int e = i * N + j, s = i * L::LD + j, t = j * L::LD + i;
          g2[p * per + e] = 0.5 * (gx2[s] + gx2[t]);
          g2[p * per + N * N + e] = 0.5 * (gy2[s] + gy2[t]);
          g1[p * per + e] = -0.5 * (gx2[s] + gx2[t]);
          g1[p * per + N * N + e] = 0.5 * (gy1[s] + gy1[t]);
        }
    }
  }
}

extern "C" int hostcheck_run(int variant, int kind, int n, int metric, int64_t b, const double* z1, const double* z2,
                             const double* w, int grad, double* dist, double* vvd, double* g1, double* g2,
                             unsigned* status) {
  if (variant == 0) {
    switch (n) {
#define CASE(K) case K: run_reg<K>(kind, metric, b, z1, z2, w, grad, dist, vvd, g1, g2, status); return 0;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
    }
    return 1;
  }
  if (variant == 2 || variant == 3) {
    if (kind != kUpper && variant == 3) return 1;
    switch (n) {
#define CASE(K)                                                                                  \
  case K:                                                                                        \
    if (variant == 2) run_coop<K, false>(kind, metric, b, z1, z2, w, grad, dist, vvd, g1, g2, status); \
    else run_coop<K, true>(kind, metric, b, z1, z2, w, grad, dist, vvd, g1, g2, status);              \
    return 0;
      CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return 1;
  }
  switch (n) {
#define CASE(K) case K: run_loc<K>(kind, metric, b, z1, z2, w, grad, dist, vvd, g1, g2, status); return 0;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
  }
  return 1;
}

// ---- optimizer rows -------------------------------------------------------------------------
#define HC_RSGD(NS)                                                                                          \
  template <int N>                                                                                           \
  static int64_t rsgd_##NS(int kind, int64_t rows, double* table, const double* grad, double lr) {           \
    constexpr int T = Cfg<N>::kTri;                                                                          \
    int64_t projected = 0;                                                                                   \
    const int per = (kind == kSpd ? 1 : 2) * N * N;                                                          \
    for (int64_t r = 0; r < rows; ++r) {                                                                     \
      double* p = table + r * per;                                                                           \
      const double* g = grad + r * per;                                                                      \
      if (kind == kSpd) {                                                                                    \
        double x[T], gx[T];                                                                                  \
        NS::pack_sym<N>(p, x);                                                                               \
        NS::pack_sym<N>(g, gx);                                                                              \
        NS::spd_rsgd_row<N>(x, gx, lr);                                                                      \
        for (int i = 0; i < N; ++i)                                                                          \
          for (int j = 0; j < N; ++j) p[i * N + j] = x[tri(i, j)];                                           \
      } else {                                                                                               \
        double x[T], y[T], gx[T], gy[T];                                                                     \
        NS::pack_sym<N>(p, x);                                                                               \
        NS::pack_sym<N>(p + N * N, y);                                                                       \
        NS::pack_sym<N>(g, gx);                                                                              \
        NS::pack_sym<N>(g + N * N, gy);                                                                      \
        projected += (kind == kBounded ? NS::bounded_rsgd_row<N>(x, y, g, g + N * N, lr)                     \
                                       : NS::upper_rsgd_row<N>(x, y, gx, gy, lr)) ? 1 : 0;                    \
        for (int i = 0; i < N; ++i)                                                                          \
          for (int j = 0; j < N; ++j) {                                                                      \
            p[i * N + j] = x[tri(i, j)];                                                                     \
            p[N * N + i * N + j] = y[tri(i, j)];                                                             \
          }                                                                                                  \
      }                                                                                                      \
    }                                                                                                        \
    return projected;                                                                                        \
  }
HC_RSGD(reg)
HC_RSGD(loc)

extern "C" int64_t hostcheck_rsgd(int variant, int kind, int n, int64_t rows, double* table, const double* grad,
                                  double lr) {
  if (variant == 0) {
    switch (n) {
#define CASE(K) case K: return rsgd_reg<K>(kind, rows, table, grad, lr);
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6)
#undef CASE
    }
    return -1;
  }
  switch (n) {
#define CASE(K) case K: return rsgd_loc<K>(kind, rows, table, grad, lr);
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
  }
  return -1;
}

// the tabulated one-directional ring schedule of the register-resident Jacobi (coop::ring_send_masks), for the
// combinatorial test in tests/test_hostcheck.py
extern "C" unsigned long long hostcheck_ring_send_masks(int g) {
  switch (g) {
    case 1: return sympa::coop::ring_send_masks<1>();
    case 2: return sympa::coop::ring_send_masks<2>();
    case 3: return sympa::coop::ring_send_masks<3>();
    case 4: return sympa::coop::ring_send_masks<4>();
    case 5: return sympa::coop::ring_send_masks<5>();
  }
  return ~0ull;
}
