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
};

template <int N>
struct CoopCfg {
  typedef coop::Layout<N> L;
  static constexpr int G = L::G;
  static constexpr int PW = 32 / G;  // pairs per warp
  static constexpr int kSmemBytes = PW * L::kDoubles * (int)sizeof(double);
};

template <int N, int KIND, int MODE>
__global__ void __launch_bounds__(32) coop_kernel(const PairArgs a) {
  static_assert(KIND == kUpper, "cooperative kernel: upper half space only");
  typedef coop::Layout<N> L;
  constexpr int G = CoopCfg<N>::G;
  constexpr int PW = CoopCfg<N>::PW;
  constexpr int NN = N * N;
  constexpr int PER = 2 * NN;
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
    WarpExec ex{g, active};
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
        const double* gx2 = sm + L::Q;
        const double* gy2 = sm + L::T2;
        const double* gy1 = sm + L::P;
        if (MODE == kModeFwdSave) {
          double* o1 = a.gz1 + p * PER;
          double* o2 = a.gz2 + p * PER;
          for (int e = g; e < NN; e += G) {
            const int et = (e % N) * N + e / N;
            const double x2 = 0.5 * (gx2[e] + gx2[et]);
            o2[e] = x2;
            o1[e] = -x2;
            o2[NN + e] = 0.5 * (gy2[e] + gy2[et]);
            o1[NN + e] = 0.5 * (gy1[e] + gy1[et]);
          }
        } else {  // fused distortion step
          const double gd = __ldg(a.graph_dist + p);
          const double r = a.scale * dist / gd;
          const double er = r * r - 1.0;
          const double sg = er > 0.0 ? 1.0 : (er < 0.0 ? -1.0 : 0.0);
          const double dl_dr = sg * 2.0 * r;
          const double dl_dd = dl_dr * a.scale / gd;
          if (g == 0) {
            loss_acc += fabs(er);
            gscale_acc += dl_dr * dist / gd;
          }
          if (a.grad_wsum_w != nullptr && a.metric == kWsum) {
            for (int k = g; k < N; k += G)
              if (a.wsum_w[k] > 0.0) atomicAdd(a.grad_wsum_w + k, dl_dd * sm[L::VS + k]);
          }
          double* o1 = a.grad_table + i1 * PER;
          double* o2 = a.grad_table + i2 * PER;
          for (int e = g; e < NN; e += G) {
            const int et = (e % N) * N + e / N;
            const double x2 = dl_dd * 0.5 * (gx2[e] + gx2[et]);
            atomicAdd(o2 + e, x2);
            atomicAdd(o1 + e, -x2);
            atomicAdd(o2 + NN + e, dl_dd * 0.5 * (gy2[e] + gy2[et]));
            atomicAdd(o1 + NN + e, dl_dd * 0.5 * (gy1[e] + gy1[et]));
          }
        }
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

template <int N, int KIND, int MODE>
static int launch_coop(const PairArgs& a, cudaStream_t s) {
  constexpr int PW = CoopCfg<N>::PW;
  constexpr int smem = CoopCfg<N>::kSmemBytes;
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
