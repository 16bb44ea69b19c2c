// One-pair-per-thread kernels: gather -> dist (-> unit gradients | -> loss + scatter-add).
// Instantiated per matrix size N in pair_kernels_n.cu (one translation unit per N so the build
// parallelises); the C ABI in sympa_b200.cu dispatches through launch_pairs<N>().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// CTA-wide re-alignment of the warps between the phases of a pair, for the sizes whose fully unrolled
// body is much larger than the instruction cache (n >= 4): see pair_kernel.
#ifndef SY_BLOCK_SYNC
#define SY_BLOCK_SYNC 1
#endif
#ifndef SY_MID_SYNC
#define SY_MID_SYNC 1
#endif
// smallest n whose CTAs are re-aligned (measured: n = 3 upper +2 %, bounded +7 %; n = 2 fits the cache)
#ifndef SY_SYNC_MIN_N
#define SY_SYNC_MIN_N 3
#endif
#if defined(SYMPA_PAIR_KERNELS_IMPL) && SY_BLOCK_SYNC && SY_MID_SYNC
template <int N>
__host__ __device__ __forceinline__ void sy_phase_sync() {
#if defined(__CUDA_ARCH__)
  if (N >= SY_SYNC_MIN_N) __syncthreads();
#endif
}
#define SY_PHASE_SYNC_REG(N) sy_phase_sync<N>();
#endif
#include "pair_math.cuh"

namespace sympa {

enum Mode { kModeFwd = 0, kModeFwdSave = 1, kModeStep = 2 };

struct PairArgs {
  int64_t num_pairs;
  const double* z1;
  const double* z2;
  const double* table;
  int64_t num_rows;
  const int64_t* idx;
  const double* wsum_w;
  double* dist_out;
  double* vvd_out;
  double* gz1;  // saved_state, first half : d dist / d z1, (num_pairs, point)
  double* gz2;  // saved_state, second half: d dist / d z2
  unsigned int* status;
  int metric;
  // fused distortion step
  const double* graph_dist;
  double scale;
  double* grad_table;
  double* grad_wsum_w;
  double* grad_scale;
  double* loss_out;
  // optional global scratch for the three-kernel split path (upper, n > SY_REG_MAX_N)
  double* scratch;
  int64_t scratch_pairs;
};

// launch the pair kernel for matrix size N; defined (explicitly instantiated) in pair_kernels_n.cu
template <int N>
int launch_pairs(int kind, int mode, const PairArgs& a, cudaStream_t stream);

struct RsgdArgs {
  int64_t num_rows;
  double* table;
  const double* grad;
  double lr;
  const double* lr_scale;           // optional device scalar multiplying lr (e.g. a clipping coefficient)
  unsigned long long* projected;    // optional device counter of rows the projection had to move
  int zero_grad;                    // != 0: the gradient rows that were applied are zeroed (fused zero_grad)
};

// launch the optimizer-row kernel for matrix size N; defined in pair_kernels_n.cu
template <int N>
int launch_rsgd(int kind, const RsgdArgs& a, cudaStream_t stream);

// Bounded domain by rows (the inverse Cayley transform acts on POINTS, not on pairs): forward maps every
// table row z -> Z = i (I + z)(I - z)^-1 once, so that the pairs can run the cheaper upper-half kernels on
// the transformed table; backward maps the accumulated upper-half gradient of a row back through the
// transform, G_z = conj(Nz) (2 G_Y - 2i G_X) conj(Nz), Nz = (I - z)^-1  (bounded_domain.py:27-39).
struct BoundedRowsArgs {
  int64_t num_rows;
  const double* table;      // (num_rows, 2, n, n) bounded points
  double* upper_out;        // forward: transformed points
  const double* grad_upper; // backward: dL/dZ rows (symmetric blocks)
  double* grad_table;       // backward: dL/dz rows
  int overwrite;            // backward: write (1) or accumulate (0)
  unsigned int* status;
};
template <int N>
int launch_bounded_rows(int backward, const BoundedRowsArgs& a, cudaStream_t stream);

int check_launch();
int grid_for(int64_t work_items, int threads, int waves_cap);

#ifdef SYMPA_PAIR_KERNELS_IMPL

#ifndef SY_PAIR_THREADS
#define SY_PAIR_THREADS 128
#endif
// waves of resident CTAs in the grid of the pair kernels (launch_one)
#ifndef SY_PAIR_WAVES
#define SY_PAIR_WAVES 8
#endif
constexpr int kThreads = SY_PAIR_THREADS;
// Re-align the warps of a CTA at the start of every pair and after its Jacobi sweeps (two
// __syncthreads per ~10k instructions): the fully unrolled body is several times larger than the
// instruction cache, and warps that walk through it together share their instruction fetches
// (measured on the B200: n = 4 forward+gradient kernel 4.06 -> 2.9 ms per 4M pairs; n = 3 fits and
// does not profit).  Every thread of the CTA therefore runs every iteration: threads without a pair
// (tail, bad index) recompute a valid one and skip the writes.
// minimum resident CTAs per SM asked of ptxas for the register-resident sizes n = 3, 4 (caps the
// registers per thread: 2 -> 255, 3 -> 168, 4 -> 128); tuned on the B200, see DESIGN.md
#ifndef SY_REG_MIN_BLOCKS
#define SY_REG_MIN_BLOCKS 2
#endif
#ifndef SY_REG_MIN_BLOCKS_3
#define SY_REG_MIN_BLOCKS_3 SY_REG_MIN_BLOCKS
#endif
// the forward-only kernel (evaluation, distance matrices) keeps less state alive: 3 CTAs/SM at 168
// registers measured +10 % (n = 4) and +7 % (n = 3) over 2 CTAs/SM (n = 5, 6 would spill: left at 2)
#ifndef SY_FWD_MIN_BLOCKS
#define SY_FWD_MIN_BLOCKS 3
#endif

// namespace switch: reg = unrolled / registers (N <= SY_REG_MAX_N), loc = rolled / local memory
template <bool REG>
struct Math;
template <>
struct Math<true> {
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double upper(A... a) { return reg::upper_pair<N, G>(a...); }
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double bounded(A... a) { return reg::bounded_pair<N, G>(a...); }
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double spd(A... a) { return reg::spd_pair<N, G>(a...); }
};
template <>
struct Math<false> {
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double upper(A... a) { return loc::upper_pair<N, G>(a...); }
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double bounded(A... a) { return loc::bounded_pair<N, G>(a...); }
  template <int N, bool G, class... A>
  static __device__ __forceinline__ double spd(A... a) { return loc::spd_pair<N, G>(a...); }
};

// one symmetric n x n block from global memory -> packed lower triangle, symmetrised
template <int N, bool REG>
__device__ __forceinline__ void load_packed(const double* __restrict__ p, double* s) {
  if (REG) {
    double buf[N * N];
    if ((N * N) % 2 == 0) {
      const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
      for (int i = 0; i < N * N / 2; ++i) {
        const double2 t = __ldg(p2 + i);
        buf[2 * i] = t.x;
        buf[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < N * N; ++i) buf[i] = __ldg(p + i);
    }
    reg::pack_sym<N>(buf, s);
  } else {
#pragma unroll 1
    for (int i = 0; i < N; ++i)
      for (int j = 0; j <= i; ++j) s[tri(i, j)] = 0.5 * (__ldg(p + i * N + j) + __ldg(p + j * N + i));
  }
}

// packed symmetric -> full n x n block in global memory
template <int N, bool REG>
__device__ __forceinline__ void store_full(double* __restrict__ out, const double* s) {
  if (REG && (N * N) % 2 == 0) {
    double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
    for (int e = 0; e < N * N / 2; ++e) {
      const int i0 = (2 * e) / N, j0 = (2 * e) % N, i1 = (2 * e + 1) / N, j1 = (2 * e + 1) % N;
      o2[e] = make_double2(s[tri(i0, j0)], s[tri(i1, j1)]);
    }
  } else if (REG) {
#pragma unroll
    for (int e = 0; e < N * N; ++e) out[e] = s[tri(e / N, e % N)];
  } else {
#pragma unroll 1
    for (int e = 0; e < N * N; ++e) out[e] = s[tri(e / N, e % N)];
  }
}

// packed lower triangles of one point's unit gradient (re, then im unless spd) -> saved state
template <int N, bool HAS_IM>
__device__ __forceinline__ void store_state(double* __restrict__ out, const double* re, const double* im) {
  constexpr int T = Cfg<N>::kTri;
  constexpr int TOT = (HAS_IM ? 2 : 1) * T;
  if (TOT % 2 == 0) {
    double2* o2 = reinterpret_cast<double2*>(out);
#pragma unroll
    for (int e = 0; e < TOT / 2; ++e) {
      const int a = 2 * e, b = 2 * e + 1;
      o2[e] = make_double2(a < T ? re[a] : im[a - T], b < T ? re[b] : im[b - T]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < TOT; ++e) out[e] = e < T ? re[e] : im[e - T];
  }
}

template <int N, bool REG>
__device__ __forceinline__ void atomic_add_full(double* __restrict__ out, const double* s, double scale) {
  if (REG) {
#pragma unroll
    for (int e = 0; e < N * N; ++e) atomicAdd(out + e, scale * s[tri(e / N, e % N)]);
  } else {
#pragma unroll 1
    for (int e = 0; e < N * N; ++e) atomicAdd(out + e, scale * s[tri(e / N, e % N)]);
  }
}


// ------------------------------------------------------------------------------------------
// Shared-memory staging of the gather and of the saved unit gradients (register kernels, n = 2..4).
//
// One pair per thread means that a per-thread row fetch touches 32 different lines per warp
// instruction (32 LDG.128 x 32 lines per pair and warp, the same again for the stores), and that the
// dependent index -> row loads sit at the head of every pair with two warps per scheduler to hide
// them: ncu attributed 19 % of the warp stall samples of pair_kernel<4> to long_scoreboard on the
// first use of the rows and 7 % to lg_throttle on the stores.  Instead, each warp copies the rows of
// its NEXT 32 pairs into shared memory with cp.async (16-byte chunks, consecutive lanes on
// consecutive chunks of a row, so an instruction touches 2-4 lines) while it computes the current
// ones, and it writes the saved unit gradients through shared memory as contiguous 512-byte
// segments (the state of 32 consecutive pairs is contiguous in HBM).  Indices are staged the same
// way two iterations ahead.  Slot strides are an odd number of 16-byte chunks: conflict-free
// LDS.128 / STS.128 when every lane walks its own slot.
#ifndef SY_STAGE_MIN_N
#define SY_STAGE_MIN_N 3
#endif
#ifndef SY_STAGE_MAX_N
#define SY_STAGE_MAX_N 4
#endif
// 1: every lane gathers the rows of its own pair (stage_rows); 0: the warp gathers row by row, coalesced
// matrix sizes whose saved state alone is staged (see StageCfg::kOutOnly).  OFF by default (max < min): measured with
// 5..6 on the B200, forward+save kernel: n = 5 2.316 -> 2.310 ms per 2^21 pairs, n = 6 4.276 -> 4.404 ms - the
// 84 STS + bulk store replace 84 STG but the kernel's queue pressure comes from its spill traffic, which stays.
#ifndef SY_OSTAGE_MIN_N
#define SY_OSTAGE_MIN_N 5
#endif
#ifndef SY_OSTAGE_MAX_N
#define SY_OSTAGE_MAX_N 0
#endif
#ifndef SY_STAGE_OWN_ROWS
#define SY_STAGE_OWN_ROWS 0
#endif

template <int N, int KIND>
struct StageCfg {
  static constexpr int PER = (KIND == kSpd ? 1 : 2) * N * N;  // doubles per point
  // (spd at n = 4 runs 3 CTAs/SM at 160 registers and measured slower with the staging: left per-thread)
  static constexpr bool kOn = (N >= SY_STAGE_MIN_N) && (N <= SY_STAGE_MAX_N) && (N <= reg_max_n(KIND)) && (PER % 2 == 0) &&
                              (KIND != kSpd);
  static constexpr int RC = PER / 2;                           // 16-byte chunks per point
  static constexpr int IN_STRIDE = 2 * RC + 1;                 // chunks per pair slot (two points + pad)
  // on the way out: the forward+save kernel (mode 1) stages the packed saved state of both operands as two
  // CONTIGUOUS blocks [operand][lane][SRC chunks] - exactly the bytes of 32 consecutive pairs in HBM, so one bulk
  // store per operand and warp moves them; the fused step (mode 2) stages one full point per lane at a time (RC
  // chunks) plus a pad chunk that carries the destination row, at an odd chunk stride, for its row-contiguous atomics
  static constexpr int SRC = state_doubles(KIND, N) / 2;
  __host__ __device__ static constexpr int out_stride(int mode) { return mode == 2 ? ((RC % 2 == 0) ? RC + 1 : RC + 2) : 2 * SRC; }
  static constexpr int IN_BYTES = 32 * IN_STRIDE * 16;
  static constexpr int IDX_BYTES = 2 * 32 * 16;                // two generations of 32 index pairs
  // per-warp layout: [rows in][indices][gradients out - not in the forward-only kernel]
  __host__ __device__ static constexpr int warp_bytes(int mode) {
    return IN_BYTES + IDX_BYTES + (mode != 0 ? 32 * out_stride(mode) * 16 : 0);
  }
  // Sizes whose rows do not fit the gather staging (n = 5, 6: 26 / 37 kB of rows per warp) still send the saved state
  // of the forward+save kernel through shared memory and one bulk store per operand and warp (15 / 21 kB per warp):
  // 84 STG.64 per thread at n = 6 otherwise, in a kernel whose load-store queue is full (lg_throttle 0.97 per issue).
  static constexpr bool kOutOnly = !kOn && (N >= SY_OSTAGE_MIN_N) && (N <= SY_OSTAGE_MAX_N) && (N <= reg_max_n(KIND)) &&
                                   (KIND != kSpd) && (state_doubles(KIND, N) % 2 == 0);
  static constexpr int OUT_ONLY_BYTES = 32 * 2 * state_doubles(KIND, N) * 8;
};

// (no "memory" clobber: the statement neither reads nor writes anything the compiler knows about before the
// cp.async.wait_all that follows it, and with the clobber every copy was a barrier for the shared-memory load
// of the NEXT copy's row index - 32 dependent LDS -> LDGSTS round trips per warp and pair, 6 % of the kernel
// time in ncu.  Ordering against the slot's previous readers is the __syncwarp() before stage_rows.)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TMA one-dimensional bulk store shared -> global (16-byte aligned addresses, bytes a multiple of 16), tracked
// by the issuing thread's bulk async-group.  A bulk copy is a warp-uniform instruction (UBLKCP): issued per lane
// the compiler serialises it in a loop over the lanes (measured: 1100 extra instructions per warp and pair),
// so it is used where ONE copy moves a warp's worth of data - the saved unit gradients of 32 consecutive
// pairs are contiguous in HBM - and not for the gather of 64 scattered rows, which stays with cp.async.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// orders this thread's generic-proxy writes to shared memory before later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// index pairs of the warp's 32 pairs starting at pair w0 -> idxb[0..31] (asynchronous).  Pairs past the
// end are clamped to the last one: every warp of a CTA that still has work computes on valid data.
__device__ __forceinline__ void stage_indices(const PairArgs& a, int64_t w0, longlong2* idxb, int lane) {
  if (a.idx == nullptr) return;
  int64_t p = w0 + lane;
  p = p < a.num_pairs ? p : a.num_pairs - 1;
  asm volatile("" ::: "memory");
  cp_async16(idxb + lane, reinterpret_cast<const longlong2*>(a.idx) + p);
  asm volatile("" ::: "memory");
}

// Owner lanes check the (landed) index pairs of a generation once, replace rejected ones by (0, 0) in
// place so that stage_rows can use them blindly, and return the warp's reject mask.
__device__ __forceinline__ unsigned validate_indices(const PairArgs& a, longlong2* idxb, int lane) {
  if (a.idx == nullptr) return 0u;
  const longlong2 ij = idxb[lane];
  const bool bad = ij.x < 0 || ij.x >= a.num_rows || ij.y < 0 || ij.y >= a.num_rows;
  if (bad) idxb[lane] = make_longlong2(0, 0);
  const unsigned m = __ballot_sync(0xffffffffu, bad);
  __syncwarp();
  return m;
}

// rows of the warp's 32 pairs starting at pair w0 -> in[] (asynchronous).  idxb holds their
// (already landed and validated) index pairs; the tail is clamped to the last pair.
template <int N, int KIND>
__device__ __forceinline__ void stage_rows(const PairArgs& a, int64_t w0, const longlong2* idxb, unsigned char* in,
                                           int lane) {
  using S = StageCfg<N, KIND>;
  // compiler fence: the copies below carry no "memory" clobber (see cp_async16), so without it the loads that
  // moved the CURRENT pairs' operands out of the slots could be scheduled after the copies that refill them
  asm volatile("" ::: "memory");
#if SY_STAGE_OWN_ROWS
  // Every lane copies the two rows of ITS OWN pair into its own slot: the row addresses sit in the lane's
  // registers (one LDS.128 of its validated index pair) and the chunk offsets are immediates - no index load
  // and no address arithmetic per 16-byte chunk.  The price: an instruction touches 32 rows instead of 2-4
  // (16 of every 32-byte sector per instruction; the other half comes with the next chunk, from L2).
  {
    const double* s1;
    const double* s2;
    if (a.idx != nullptr) {
      const longlong2 ij = idxb[lane];
      s1 = a.table + ij.x * S::PER;
      s2 = a.table + ij.y * S::PER;
    } else {
      int64_t p = w0 + lane;
      p = p < a.num_pairs ? p : a.num_pairs - 1;
      s1 = a.z1 + p * S::PER;
      s2 = a.z2 + p * S::PER;
    }
    unsigned char* slot = in + lane * (S::IN_STRIDE * 16);
#pragma unroll
    for (int c = 0; c < S::RC; ++c) {
      cp_async16(slot + c * 16, s1 + 2 * c);
      cp_async16(slot + (S::RC + c) * 16, s2 + 2 * c);
    }
    asm volatile("" ::: "memory");
    return;
  }
#endif
  // flat chunk f = it * 32 + lane among the warp's 32 pairs: pair j = f / (2 RC), then the point
  // (side 0 / 1) and the chunk c within the point
  if (a.idx != nullptr) {
    const int64_t* rows = reinterpret_cast<const int64_t*>(idxb);  // validated by validate_indices
#pragma unroll
    for (int it = 0; it < 2 * S::RC; ++it) {
      const int f = it * 32 + lane;
      const int j = f / (2 * S::RC);
      const int rem = f - j * (2 * S::RC);
      const int side = rem / S::RC;
      const int c = rem - side * S::RC;
      cp_async16(in + (j * S::IN_STRIDE + rem) * 16, a.table + rows[2 * j + side] * S::PER + 2 * c);
    }
  } else {
#pragma unroll
    for (int it = 0; it < 2 * S::RC; ++it) {
      const int f = it * 32 + lane;
      const int j = f / (2 * S::RC);
      const int rem = f - j * (2 * S::RC);
      const int side = rem / S::RC;
      const int c = rem - side * S::RC;
      int64_t p = w0 + j;
      p = p < a.num_pairs ? p : a.num_pairs - 1;
      cp_async16(in + (j * S::IN_STRIDE + rem) * 16, (side ? a.z2 : a.z1) + p * S::PER + 2 * c);
    }
  }
  asm volatile("" ::: "memory");
}

// one symmetric n x n block from shared memory -> packed lower triangle, symmetrised
template <int N>
__device__ __forceinline__ void load_packed_smem(const double* p, double* s) {
  double buf[N * N];
  if ((N * N) % 2 == 0) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < N * N / 2; ++i) {
      const double2 t = p2[i];
      buf[2 * i] = t.x;
      buf[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N * N; ++i) buf[i] = p[i];
  }
  reg::pack_sym<N>(buf, s);
}

// the warp's 32 staged (already scaled) point gradients -> scatter-add into the table gradient,
// consecutive lanes on consecutive elements of a row (one 32-byte sector per four lanes instead of
// one per lane).  The pad chunk of a slot holds the destination row, or -1 for "no contribution".
template <int N, int KIND>
__device__ __forceinline__ void flush_atomic(const unsigned char* out, double* __restrict__ grad_table, int lane) {
  using S = StageCfg<N, KIND>;
  constexpr int STRIDE = S::out_stride(2);
#pragma unroll
  for (int it = 0; it < S::PER; ++it) {
    const int f = it * 32 + lane;
    const int j = f / S::PER;
    const int e = f - j * S::PER;
    const int64_t r = *reinterpret_cast<const int64_t*>(out + (j * STRIDE + S::RC) * 16);
    if (r >= 0) atomicAdd(grad_table + r * S::PER + e, *reinterpret_cast<const double*>(out + j * STRIDE * 16 + e * 8));
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int N, int KIND, int MODE>
__global__ void __launch_bounds__(kThreads, (N <= reg_max_n(KIND) && N >= 3)
                                                ? ((MODE == 0 && N <= 4) ? SY_FWD_MIN_BLOCKS
                                                   : (N == 3 && KIND == 1 && MODE == 1) ? 3  /* bounded n = 3 fwd+save: +6 % at 3 CTAs/SM */
                                                   : (N == 3 ? SY_REG_MIN_BLOCKS_3 : SY_REG_MIN_BLOCKS))
                                                : 1)
    pair_kernel(const PairArgs a) {
  constexpr bool REG = N <= reg_max_n(KIND);
  constexpr int T = Cfg<N>::kTri;
  constexpr int PER = (KIND == kSpd ? 1 : 2) * N * N;
  constexpr bool GRAD = MODE != kModeFwd;
  using M = Math<REG>;
  unsigned st = 0;
  double loss_acc = 0.0, gscale_acc = 0.0;
  double gw_acc[(MODE == kModeStep) ? N : 1];
  if (MODE == kModeStep) {
#pragma unroll
    for (int k = 0; k < N; ++k) gw_acc[k] = 0.0;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  using S = StageCfg<N, KIND>;
  constexpr bool STAGE = S::kOn;
  const int lane = threadIdx.x & 31;
  const int wofs = (int)(threadIdx.x & ~31u);  // first pair of this warp within the CTA's batch
  unsigned char* st_in = nullptr;              // [32 pair slots]   rows of the current / next pairs
  unsigned char* st_out = nullptr;             // [32 point slots]  unit gradients on their way out
  longlong2* st_idx = nullptr;                 // [2][32]           index pairs, two generations
  unsigned bad_next = 0u;                      // lanes whose staged (next) index pair was rejected
  constexpr bool OSTAGE = S::kOutOnly && MODE == kModeFwdSave;
  if (OSTAGE) {
    extern __shared__ __align__(16) unsigned char sy_stage_smem[];
    st_out = sy_stage_smem + (threadIdx.x >> 5) * S::OUT_ONLY_BYTES;
  }
  if (STAGE) {
    extern __shared__ __align__(16) unsigned char sy_stage_smem[];
    unsigned char* wb = sy_stage_smem + (threadIdx.x >> 5) * S::warp_bytes(MODE);
    st_in = wb;
    st_idx = reinterpret_cast<longlong2*>(wb + S::IN_BYTES);
    st_out = wb + S::IN_BYTES + S::IDX_BYTES;
    const int64_t b0 = (int64_t)blockIdx.x * blockDim.x;
    if (b0 < a.num_pairs) {
      stage_indices(a, b0 + wofs, st_idx, lane);
      if (b0 + stride < a.num_pairs) stage_indices(a, b0 + stride + wofs, st_idx + 32, lane);
      cp_async_wait_all();
      __syncwarp();
      bad_next = validate_indices(a, st_idx, lane);
      stage_rows<N, KIND>(a, b0 + wofs, st_idx, st_in, lane);
      cp_async_wait_all();   // the first rows: exposed once per CTA
    }
  }
#ifdef SY_DEPHASE_NS
  if (blockIdx.x & 1) __nanosleep(SY_DEPHASE_NS);   // experiment: start the odd CTAs half an iteration late
#endif
  int gen = 0;  // generation (parity) of the index buffer holding the current pairs
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < a.num_pairs; base += stride, gen ^= 1) {
    if (SY_BLOCK_SYNC && REG && N >= SY_SYNC_MIN_N) __syncthreads();
    int64_t p = base + threadIdx.x;
    bool active = p < a.num_pairs;
    if (!active) p = a.num_pairs - 1;
    const double* p1;
    const double* p2;
    int64_t i1 = 0, i2 = 0;
    if (STAGE) {
      // (the rows of this iteration and the indices of the next were waited for at the END of the previous
      // iteration, before its bulk stores were issued: LDGSTS groups and bulk groups are tracked by the same
      // dependency barrier, and a wait placed here also sat out the stores just issued - 0.6 us per iteration)
      __syncwarp();
      if (a.idx != nullptr) {
        if (MODE == kModeStep) {  // only the fused step scatters by row
          const longlong2 ij = st_idx[gen * 32 + lane];
          i1 = ij.x;
          i2 = ij.y;
        }
        if ((bad_next >> lane) & 1u) {
          if (active) {
            st |= kStatusBadIndex;
            if (a.dist_out) a.dist_out[p] = 0.0;
          }
          active = false;
        }
      }
      p1 = reinterpret_cast<const double*>(st_in + lane * (S::IN_STRIDE * 16));
      p2 = p1 + S::PER;
    } else if (a.idx != nullptr) {
      const longlong2 ij = __ldg(reinterpret_cast<const longlong2*>(a.idx) + p);
      i1 = ij.x;
      i2 = ij.y;
      if (i1 < 0 || i1 >= a.num_rows || i2 < 0 || i2 >= a.num_rows) {
        if (active) {
          st |= kStatusBadIndex;
          if (a.dist_out) a.dist_out[p] = 0.0;
        }
        active = false;
        i1 = 0;
        i2 = 0;
      }
      p1 = a.table + i1 * PER;
      p2 = a.table + i2 * PER;
    } else {
      p1 = a.z1 + p * PER;
      p2 = a.z2 + p * PER;
    }
    if (!STAGE) {  // L2 prefetch of the rows of this thread's next pair
      const int64_t pn = base + stride + threadIdx.x;
      if (pn < a.num_pairs) {
        const double* q1;
        const double* q2;
        bool okn = true;
        if (a.idx != nullptr) {
          const longlong2 jn = __ldg(reinterpret_cast<const longlong2*>(a.idx) + pn);
          okn = jn.x >= 0 && jn.x < a.num_rows && jn.y >= 0 && jn.y < a.num_rows;
          q1 = a.table + (okn ? jn.x : 0) * PER;
          q2 = a.table + (okn ? jn.y : 0) * PER;
        } else {
          q1 = a.z1 + pn * PER;
          q2 = a.z2 + pn * PER;
        }
        if (okn) {
#pragma unroll
          for (int off = 0; off < PER; off += 16) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q1 + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q2 + off));
          }
        }
      }
    }
    double vs[N];
    double dist;
    unsigned stp = 0;  // status of this pair; dropped when the thread only recomputes a stand-in pair
    double g1r[GRAD ? T : 1], g1i[GRAD ? T : 1], g2r[GRAD ? T : 1], g2i[GRAD ? T : 1];
    if (KIND == kSpd) {
      double x[T], y[T];
      if (STAGE) {
        load_packed_smem<N>(p1, x);
        load_packed_smem<N>(p2, y);
        __syncwarp();
        if (base + 2 * stride < a.num_pairs) stage_indices(a, base + 2 * stride + wofs, st_idx + gen * 32, lane);
        if (base + stride < a.num_pairs) {
          bad_next = validate_indices(a, st_idx + (gen ^ 1) * 32, lane);
          stage_rows<N, KIND>(a, base + stride + wofs, st_idx + (gen ^ 1) * 32, st_in, lane);
        }
      } else {
        load_packed<N, REG>(p1, x);
        load_packed<N, REG>(p2, y);
      }
      dist = M::template spd<N, GRAD>(x, y, vs, g1r, g2r, &stp);
    } else {
      double x1[T], y1[T], x2[T], y2[T];
      if (STAGE) {
        load_packed_smem<N>(p1, x1);
        load_packed_smem<N>(p1 + N * N, y1);
        load_packed_smem<N>(p2, x2);
        load_packed_smem<N>(p2 + N * N, y2);
        // every lane has its operands in registers: refill the slots with the next pairs' rows
        __syncwarp();
        if (base + 2 * stride < a.num_pairs) stage_indices(a, base + 2 * stride + wofs, st_idx + gen * 32, lane);
        if (base + stride < a.num_pairs) {
          bad_next = validate_indices(a, st_idx + (gen ^ 1) * 32, lane);
          stage_rows<N, KIND>(a, base + stride + wofs, st_idx + (gen ^ 1) * 32, st_in, lane);
        }
      } else {
        load_packed<N, REG>(p1, x1);
        load_packed<N, REG>(p1 + N * N, y1);
        load_packed<N, REG>(p2, x2);
        load_packed<N, REG>(p2 + N * N, y2);
      }
      if (KIND == kUpper)
        dist = M::template upper<N, GRAD>(x1, y1, x2, y2, a.metric, a.wsum_w, vs, g1r, g1i, g2r, g2i, &stp);
      else
        dist = M::template bounded<N, GRAD>(x1, y1, x2, y2, a.metric, a.wsum_w, vs, g1r, g1i, g2r, g2i, &stp);
    }
    if (active) st |= stp;
    if (active && a.dist_out) a.dist_out[p] = dist;
    if (active && a.vvd_out) {
#pragma unroll
      for (int k = 0; k < N; ++k) a.vvd_out[p * N + k] = vs[k];
    }
    if (STAGE) cp_async_wait_all();   // next iteration's rows and indices (issued at the top of this one): long done
    if ((STAGE || OSTAGE) && MODE == kModeFwdSave) {
      // unit gradients -> the warp's two contiguous blocks -> one bulk store per operand into the saved state
      // (zeros for a pair whose indices were rejected; the tail is cut off by the byte count)
      constexpr int PS = state_doubles(KIND, N);
      double* o1 = reinterpret_cast<double*>(st_out) + lane * PS;
      double* o2 = o1 + 32 * PS;
      if (!active) {
#pragma unroll
        for (int i = 0; i < T; ++i) {
          g1r[i] = 0.0;
          g2r[i] = 0.0;
          if (KIND != kSpd) {
            g1i[i] = 0.0;
            g2i[i] = 0.0;
          }
        }
      }
      if (lane == 0) bulk_wait_read();   // the stores of the previous iteration have read the blocks (long ago)
      __syncwarp();
      store_state<N, KIND != kSpd>(o1, g1r, g1i);
      store_state<N, KIND != kSpd>(o2, g2r, g2i);
      fence_async_smem();
      __syncwarp();
      const int64_t w0 = base + wofs;
      const int64_t left = a.num_pairs - w0;   // pairs of this warp that exist (<= 0: a warp of the tail)
      if (lane == 0 && left > 0) {
        const unsigned bytes = (unsigned)((left < 32 ? left : 32) * (PS * 8));
        bulk_s2g(a.gz1 + w0 * PS, st_out, bytes);
        bulk_s2g(a.gz2 + w0 * PS, reinterpret_cast<double*>(st_out) + 32 * PS, bytes);
        bulk_commit();
      }
    } else if (active && MODE == kModeFwdSave) {
      constexpr int PS = state_doubles(KIND, N);   // packed saved state
      store_state<N, KIND != kSpd>(a.gz1 + p * PS, g1r, g1i);
      store_state<N, KIND != kSpd>(a.gz2 + p * PS, g2r, g2i);
    }
    if (active && MODE == kModeStep) {
      // L_p = |(s d / g)^2 - 1|   (sympa/losses.py:16-19 with the scale of sympa/model.py:30)
      const double gd = __ldg(a.graph_dist + p);
      const double r = a.scale * dist / gd;
      const double e = r * r - 1.0;
      loss_acc += fabs(e);
      const double sg = e > 0.0 ? 1.0 : (e < 0.0 ? -1.0 : 0.0);  // torch.abs backward: sign, 0 at 0
      const double dl_dr = sg * 2.0 * r;
      const double dl_dd = dl_dr * a.scale / gd;
      gscale_acc += dl_dr * dist / gd;
      if (a.grad_wsum_w != nullptr && a.metric == kWsum) {
#pragma unroll
        for (int k = 0; k < N; ++k) gw_acc[k] += (a.wsum_w[k] > 0.0) ? dl_dd * vs[k] : 0.0;
      }
      if (STAGE) {
#pragma unroll
        for (int i = 0; i < T; ++i) {
          g1r[i] *= dl_dd;
          g2r[i] *= dl_dd;
          if (KIND != kSpd) {
            g1i[i] *= dl_dd;
            g2i[i] *= dl_dd;
          }
        }
      } else {
        double* o1 = a.grad_table + i1 * PER;
        double* o2 = a.grad_table + i2 * PER;
        atomic_add_full<N, REG>(o1, g1r, dl_dd);
        atomic_add_full<N, REG>(o2, g2r, dl_dd);
        if (KIND != kSpd) {
          atomic_add_full<N, REG>(o1 + N * N, g1i, dl_dd);
          atomic_add_full<N, REG>(o2 + N * N, g2i, dl_dd);
        }
      }
    }
    if (STAGE && MODE == kModeStep) {
      // scaled gradients -> own slot (+ destination row in the pad chunk) -> row-contiguous atomics
      double* os = reinterpret_cast<double*>(st_out + lane * (S::out_stride(2) * 16));
      int64_t* orow = reinterpret_cast<int64_t*>(st_out + (lane * S::out_stride(2) + S::RC) * 16);
      store_full<N, REG>(os, g1r);
      if (KIND != kSpd) store_full<N, REG>(os + N * N, g1i);
      *orow = active ? i1 : -1;
      __syncwarp();
      flush_atomic<N, KIND>(st_out, a.grad_table, lane);
      __syncwarp();
      store_full<N, REG>(os, g2r);
      if (KIND != kSpd) store_full<N, REG>(os + N * N, g2i);
      *orow = active ? i2 : -1;
      __syncwarp();
      flush_atomic<N, KIND>(st_out, a.grad_table, lane);
    }
  }
  if (STAGE) cp_async_wait_all();
  if ((STAGE || OSTAGE) && MODE == kModeFwdSave && lane == 0) bulk_wait_all();
  if (MODE == kModeStep) {
    loss_acc = warp_sum(loss_acc);
    gscale_acc = warp_sum(gscale_acc);
    if ((threadIdx.x & 31) == 0) {
      if (a.loss_out) atomicAdd(a.loss_out, loss_acc);
      if (a.grad_scale) atomicAdd(a.grad_scale, gscale_acc);
    }
    if (a.grad_wsum_w != nullptr && a.metric == kWsum) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double v = warp_sum(gw_acc[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.grad_wsum_w + k, v);
      }
    }
  }
  if (st != 0 && a.status != nullptr) atomicOr(a.status, st);
}

template <int N, int KIND, int MODE>
static int launch_one(const PairArgs& a, cudaStream_t s) {
  int smem = 0;
  if (StageCfg<N, KIND>::kOn) smem = (kThreads / 32) * StageCfg<N, KIND>::warp_bytes(MODE);
  if (StageCfg<N, KIND>::kOutOnly && MODE == kModeFwdSave) smem = (kThreads / 32) * StageCfg<N, KIND>::OUT_ONLY_BYTES;
  static bool configured[64] = {};  // per instantiation and device (the attribute is per function per device)
  static int resident[64] = {};     // CTAs of this instantiation that fit on one SM
  int dev = 0;
  cudaGetDevice(&dev);
  int occ = 0;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (smem > 0 &&
        cudaFuncSetAttribute(pair_kernel<N, KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return check_launch();
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pair_kernel<N, KIND, MODE>, kThreads, smem) != cudaSuccess || occ < 1)
      occ = 1;
    if (dev >= 0 && dev < 64) {
      configured[dev] = true;
      resident[dev] = occ;
    }
  } else {
    occ = resident[dev];
  }
  // Grid-stride launch of SY_PAIR_WAVES x (the CTAs resident at once).  Measured at n = 4, 2^23 pairs: 1 wave
  // (fully persistent) 3.67 ms, 2 waves 3.58 ms, 8 waves 3.44 ms - with a single wave the two CTAs of an SM run
  // in lockstep, both in their gather / store phases at the same time, and the FP64 pipe idles; CTAs that
  // retire and start at different times de-phase the SM.
  const int grid = grid_for(a.num_pairs, kThreads, SY_PAIR_WAVES * occ);
  pair_kernel<N, KIND, MODE><<<grid, kThreads, smem, s>>>(a);
  return check_launch();
}

template <int N, int KIND, int MODE>
static int launch_coop(const PairArgs& a, cudaStream_t s);  // coop_kernels.cuh
template <int N, int MODE>
static int launch_split(const PairArgs& a, double* scratch, int64_t cap, cudaStream_t s);

// kernel selection: one pair per thread in registers for n <= reg_max_n(kind); warp-cooperative shared
// memory kernel for the larger upper-half sizes; rolled per-thread fallback otherwise.
template <int N, int KIND, int MODE>
static int launch_any(const PairArgs& a, cudaStream_t s) {
  if constexpr (KIND == kUpper && (N > reg_max_n(kUpper))) {
    if (a.scratch != nullptr && a.scratch_pairs > 0) return launch_split<N, MODE>(a, a.scratch, a.scratch_pairs, s);
    return launch_coop<N, KIND, MODE>(a, s);
  } else if constexpr (N > reg_max_n(KIND)) {   // spd, bounded
    return launch_coop<N, KIND, MODE>(a, s);
  } else {
    return launch_one<N, KIND, MODE>(a, s);
  }
}

template <int N, int KIND>
static int launch_mode(int mode, const PairArgs& a, cudaStream_t s) {
  switch (mode) {
    case kModeFwd: return launch_any<N, KIND, kModeFwd>(a, s);
    case kModeFwdSave: return launch_any<N, KIND, kModeFwdSave>(a, s);
    case kModeStep: return launch_any<N, KIND, kModeStep>(a, s);
  }
  return 1;
}

// Riemannian SGD row update (egrad2rgrad + retr + projx fused; upper_half.py:25-66, geoopt RSGD):
// one thread per table row, rows whose gradient is identically zero are left untouched - which makes
// the dense launch equivalent to a sparse update of the rows the batch touched.
// largest matrix size whose optimizer rows run the unrolled (register) templates; above it the rolled
// local-memory ones
#ifndef SY_RSGD_REG_MAX_N
#define SY_RSGD_REG_MAX_N 10
#endif
template <int N, int KIND>
__global__ void __launch_bounds__(kThreads) rsgd_kernel(const RsgdArgs a) {
  // (bounded: the projection carries a full complex Jacobi - unrolled at n = 10 it is a 32 kB stack frame)
  constexpr bool REG = N <= (KIND == kBounded ? reg_max_n(kBounded) : SY_RSGD_REG_MAX_N);
  constexpr int T = Cfg<N>::kTri;
  constexpr int PER = (KIND == kSpd ? 1 : 2) * N * N;
  const double lr = a.lr * (a.lr_scale != nullptr ? __ldg(a.lr_scale) : 1.0);
  unsigned long long moved = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.num_rows; r += stride) {
    const double* g = a.grad + r * PER;
    double* p = a.table + r * PER;
    double x[T], y[T];
    if (KIND == kBounded) {
      // the bounded update needs the gradient block as it is (bounded_rsgd_row): full, not symmetrised
      double gf[PER];
      bool any = false;
      if (REG) {
#pragma unroll
        for (int e = 0; e < PER; ++e) gf[e] = __ldg(g + e);
#pragma unroll
        for (int e = 0; e < PER; ++e) any = any || (gf[e] != 0.0);
      } else {
#pragma unroll 1
        for (int e = 0; e < PER; ++e) {
          gf[e] = __ldg(g + e);
          any = any || (gf[e] != 0.0);
        }
      }
      if (!any) continue;
      if (a.zero_grad) {
        double* gw = const_cast<double*>(g);
#pragma unroll 4
        for (int e = 0; e < PER; ++e) gw[e] = 0.0;
      }
      load_packed<N, REG>(p, x);
      load_packed<N, REG>(p + N * N, y);
      const bool m = REG ? reg::bounded_rsgd_row<N>(x, y, gf, gf + N * N, lr) : loc::bounded_rsgd_row<N>(x, y, gf, gf + N * N, lr);
      moved += m ? 1ull : 0ull;
      store_full<N, REG>(p, x);
      store_full<N, REG>(p + N * N, y);
      continue;
    }
    double gx[T], gy[T];
    load_packed<N, REG>(g, gx);
    if (KIND != kSpd) load_packed<N, REG>(g + N * N, gy);
    bool any = false;
#pragma unroll
    for (int i = 0; i < T; ++i) any = any || (gx[i] != 0.0) || (KIND != kSpd && gy[i] != 0.0);
    if (!any) continue;
    if (a.zero_grad) {  // the caller owns the gradient buffer and asked for optimizer.zero_grad() in the same pass
      double* gw = const_cast<double*>(g);
#pragma unroll 4
      for (int e = 0; e < PER; ++e) gw[e] = 0.0;
    }
    load_packed<N, REG>(p, x);
    if (KIND == kSpd) {
      if (REG) reg::spd_rsgd_row<N>(x, gx, lr); else loc::spd_rsgd_row<N>(x, gx, lr);
      store_full<N, REG>(p, x);
    } else {
      load_packed<N, REG>(p + N * N, y);
      const bool m = REG ? reg::upper_rsgd_row<N>(x, y, gx, gy, lr) : loc::upper_rsgd_row<N>(x, y, gx, gy, lr);
      moved += m ? 1ull : 0ull;
      store_full<N, REG>(p, x);
      store_full<N, REG>(p + N * N, y);
    }
  }
  if (a.projected != nullptr && moved != 0) atomicAdd(a.projected, moved);
}

template <int N>
int launch_rsgd(int kind, const RsgdArgs& a, cudaStream_t s) {
  const int grid = grid_for(a.num_rows, kThreads, 16);
  if (kind == kUpper)
    rsgd_kernel<N, kUpper><<<grid, kThreads, 0, s>>>(a);
  else if (kind == kSpd)
    rsgd_kernel<N, kSpd><<<grid, kThreads, 0, s>>>(a);
  else if (kind == kBounded)
    rsgd_kernel<N, kBounded><<<grid, kThreads, 0, s>>>(a);
  else
    return 2;
  return check_launch();
}

// largest matrix size whose row transforms run the unrolled templates (round 1 stopped at 7 and left the larger
// bounded sizes on the per-pair cooperative kernel with its 12 shared-memory buffers; unrolled, the rows of n = 10
// cost two real SPD inversions and one complex sandwich per ROW, and the pairs run the upper-half kernel)
#ifndef SY_BOUNDED_ROWS_REG_MAX_N
#define SY_BOUNDED_ROWS_REG_MAX_N 10
#endif
template <int N, bool BACKWARD>
__global__ void __launch_bounds__(kThreads) bounded_rows_kernel(const BoundedRowsArgs a) {
  constexpr bool REG = N <= SY_BOUNDED_ROWS_REG_MAX_N;
  constexpr int T = Cfg<N>::kTri;
  constexpr int PER = 2 * N * N;
  unsigned st = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.num_rows; r += stride) {
    double zr[T], zi[T], x[T], y[T], u[T], v[T];
    load_packed<N, REG>(a.table + r * PER, zr);
    load_packed<N, REG>(a.table + r * PER + N * N, zi);
    const bool ok = REG ? reg::bounded_to_upper<N>(zr, zi, x, y, u, v) : loc::bounded_to_upper<N>(zr, zi, x, y, u, v);
    if (!BACKWARD) {
      if (!ok) st |= kStatusNotPD;
      store_full<N, REG>(a.upper_out + r * PER, x);
      store_full<N, REG>(a.upper_out + r * PER + N * N, y);
    } else {
      double gx[T], gy[T], nr[T], ni[T], gr[T], gi[T];
      load_packed<N, REG>(a.grad_upper + r * PER, gx);
      load_packed<N, REG>(a.grad_upper + r * PER + N * N, gy);
#pragma unroll
      for (int i = 0; i < T; ++i) {
        nr[i] = 2.0 * gy[i];
        ni[i] = -2.0 * gx[i];
        v[i] = -v[i];
      }
      if (REG) reg::csym_sandwich<N>(u, v, nr, ni, 1.0, gr, gi); else loc::csym_sandwich<N>(u, v, nr, ni, 1.0, gr, gi);
      double* o = a.grad_table + r * PER;
      if (!a.overwrite) {
        double pr[T], pi[T];
        load_packed<N, REG>(o, pr);
        load_packed<N, REG>(o + N * N, pi);
#pragma unroll
        for (int i = 0; i < T; ++i) {
          gr[i] += pr[i];
          gi[i] += pi[i];
        }
      }
      store_full<N, REG>(o, gr);
      store_full<N, REG>(o + N * N, gi);
    }
  }
  if (!BACKWARD && st != 0 && a.status != nullptr) atomicOr(a.status, st);
}

template <int N>
int launch_bounded_rows(int backward, const BoundedRowsArgs& a, cudaStream_t s) {
  const int grid = grid_for(a.num_rows, kThreads, 16);
  if (backward)
    bounded_rows_kernel<N, true><<<grid, kThreads, 0, s>>>(a);
  else
    bounded_rows_kernel<N, false><<<grid, kThreads, 0, s>>>(a);
  return check_launch();
}

template <int N>
int launch_pairs(int kind, int mode, const PairArgs& a, cudaStream_t s) {
  switch (kind) {
    case kUpper: return launch_mode<N, kUpper>(mode, a, s);
    case kBounded: return launch_mode<N, kBounded>(mode, a, s);
    case kSpd: return launch_mode<N, kSpd>(mode, a, s);
  }
  return 2;
}

#endif  // SYMPA_PAIR_KERNELS_IMPL

}  // namespace sympa
