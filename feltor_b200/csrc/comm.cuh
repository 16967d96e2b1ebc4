// Internal: NCCL-backed communicator for the y-decomposed (slab) solver.
#pragma once
#include "common.cuh"
namespace dgb {
struct Comm;
int comm_rank(const Comm* c);
int comm_size(const Comm* c);
// in-place sum over ranks of `count` int64 words on `st`
int comm_allreduce_i64(Comm* c, long long* buf, size_t count, cudaStream_t st);
// exchange `ghost_rows` rows of `row_len` doubles with the lower / upper neighbour of a ring (periodic) or chain:
// interior = first interior row of a buffer laid out [ghost_rows | nrows | ghost_rows]
int comm_halo_rows(Comm* c, double* interior, size_t row_len, size_t nrows, size_t ghost_rows, int periodic, cudaStream_t st);
}  // namespace dgb
