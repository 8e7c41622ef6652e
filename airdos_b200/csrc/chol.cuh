// chol.cuh -- dense FP64 Cholesky solve in one cluster launch (chol.cu); shared with ba.cu.
#pragma once
#include "common.cuh"

namespace adb {

constexpr int kCholNB = 32;        // tile order
constexpr int kCholThreads = 256;  // 8 warps per CTA (255 registers per thread: the tile factorisation keeps an 8 x 8 block in registers)

// Padded layout of the system array: ld = nblk * 32 columns, (nblk + 1) * 32 rows (row ld = right-hand side).
inline int chol_nblk(int n) { return (n + kCholNB - 1) / kCholNB; }
inline size_t chol_elems(int n) { const size_t ld = (size_t)chol_nblk(n) * kCholNB; return (ld + kCholNB) * ld; }

// Factors S in place (lower triangle, right-hand side in row ld) and writes the solution (ld entries, pad = 0) to x.
// scratch: chol_scratch_elems(n) doubles (inverted / published diagonal factors).  cluster: 8 / 16, or 0 = choose by order.  *info must be zero on entry.
size_t chol_scratch_elems(int n);
adb_status chol_solve_launch(cudaStream_t st, double* S, int ld, int nblk, double* scratch, double* x, int* info, int cluster, const int* skip = nullptr,
                             long long* prof = nullptr);   // *skip != 0 (device memory) turns the launch into a no-op
int chol_max_cluster();

}  // namespace adb
