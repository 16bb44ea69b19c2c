// One translation unit per matrix size: compile with -DSYMPA_TU_N=<n> (see __graft_entry__.build).
#ifndef SYMPA_TU_N
#error "compile with -DSYMPA_TU_N=<n>"
#endif
#define SYMPA_PAIR_KERNELS_IMPL
#include "pair_kernels.cuh"
#include "coop_kernels.cuh"

namespace sympa {
template int launch_pairs<SYMPA_TU_N>(int, int, const PairArgs&, cudaStream_t);
template int launch_rsgd<SYMPA_TU_N>(int, const RsgdArgs&, cudaStream_t);
template int launch_bounded_rows<SYMPA_TU_N>(int, const BoundedRowsArgs&, cudaStream_t);
}
