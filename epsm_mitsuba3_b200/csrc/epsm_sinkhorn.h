// epsm_sinkhorn.h — launch wrappers of the Sinkhorn matcher kernels (epsm_sinkhorn.cu); internal, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

// One softmin of a Sinkhorn step: pot_rows_new[i] = softmin_j(ε; C(rows_i, cols_j), w_log + pot_cols_j/ε).
struct SinkJob {
    const float *rows, *cols;          // n_rows × 5, n_cols × 5
    const float *pot_cols;             // potential carried by the columns (NULL = 0: the initialisation step)
    const float *pot_rows_old;         // averaged with when SinkStep::average (NULL = never)
    float *pot_rows_new;
    float *bary;                       // n_rows × 5: Σ_j w_ij cols_j (gradient launch only; may be NULL)
    uint32_t n_rows, n_cols;
    float w_log;                       // ln of the uniform column weight
    uint32_t _pad;
    float4 *packed;                    // scratch of 6·ceil(n_cols/4) float4: the columns as k_sink_pack lays them out for this step
    float *tile_info;                  // scratch of 12 floats per 1024-column tile: box of its points and largest h·log2 e
};
inline size_t epsm_sink_tile_floats(uint32_t n_cols) { return 12 * (size_t) ((n_cols + 1023) / 1024); }
inline size_t epsm_sink_packed_floats(uint32_t n_cols) { return 24 * (size_t) ((n_cols + 3) / 4); }
struct SinkStep { SinkJob job[4]; float eps; int average; };

cudaError_t epsm_launch_sink_bbox(const float *x, uint32_t n, const float *y, uint32_t m, int *out10, cudaStream_t s);
float epsm_sink_ord_to_float(int ordered);
cudaError_t epsm_launch_sink_step(const SinkStep &st, int n_jobs, bool grad, cudaStream_t s);
size_t epsm_sink_order_temp_bytes(uint32_t n);
cudaError_t epsm_launch_sink_order(const float *pts, uint32_t n, float lo3, float hi3, float lo4, float hi4, uint32_t *words, void *temp, size_t temp_bytes,
                                   float *sorted, cudaStream_t s);
cudaError_t epsm_launch_sink_finish(const float *bary_xx, const float *bary_xy, float factor, uint32_t n, float *grad, const uint32_t *perm,
                                    const float *f_ba, const float *f_aa, const float *g_ab, const float *g_bb, uint32_t m, double *loss,
                                    cudaStream_t s);
