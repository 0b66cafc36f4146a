// tail_kernel.cuh — parameters of the fused encoder tail (tail_kernel.cu): y = normalize_rows(x . W).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "clip_kernels.cuh"

namespace flyp {

struct TailParams {
    int n;              // rows of x
    int n_out;          // N: output columns (multiple of 64, <= 1024)
    int n_cta;          // columns computed by one CTA (multiple of 64, <= 512): N, or N / 2 rounded up to 64 on a CTA pair
    int kc;             // 64-wide K chunks per plane
    KPlan kplan;        // bf16: one term; fp32 inputs: six products of the 3-way bf16 splits
    int w_plane_rows;   // row offset between the planes of W ([3][K][N] stacked)
    int stages;         // ring depth (set by launch_tail)
    void* y;            // [n][N] bf16 or fp32
    int y_fp32;
    void* y16;          // optional [n][N] fp16 copy of the rounded bf16 features
    float* inv_norm;    // optional [n]: 1 / ||z||
};

// CTAs that share a row tile (1, or 2 for N > 512)
int tail_n_split(int n_out);
// tmX: x (or its bf16 planes) with [128 rows][64 k] boxes; tmW: W (or its planes, stacked by rows) with [64 k][64 n] boxes
void launch_tail(const CUtensorMap& tmX, const CUtensorMap& tmW, TailParams p, cudaStream_t st);

}  // namespace flyp
