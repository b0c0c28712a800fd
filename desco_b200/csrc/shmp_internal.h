// Internal interface between shmp.cu (C-ABI entry points, layered fp32 path, readout MLPs) and shmp_fused.cu
// (the fused tcgen05 layer kernel).  Not part of the public C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SHMP_TILE_ROWS 128     /* rows per fused tile = one UMMA M                                   */
#define SHMP_TILE_MAX_NBH 20   /* neighborhoods per fused tile (bounds the canonical-row buffers)    */
#define SHMP_PLAN_CHUNK 2048   /* neighborhoods per tile-planning CTA; tiles never straddle chunks   */
/* bytes of one layer in the tensor-core weight blob: B hi/lo images (2 x 192 x 128), bias_c, bias_a (64 fp32 each),
 * Wa [192][64] and Cw [64][128] as mma.sync B fragments (bf16 hi/lo, tcpack.pack_mma_b_frags; same bytes as fp32) */
#define SHMP_TC_LAYER_BYTES (2 * 192 * 128 + 2 * 64 * 4 + 64 * 192 * 4 + 128 * 64 * 4)

int64_t desco_internal_shmp_fused_workspace_bytes(int num_neighborhoods);

// Runs every message-passing layer for a hetero (count/canonical) batch and leaves the per-layer pooled count rows
// and canonical rows in pool / emb_a ([G][emb_ld], emb_ld = (layers + 1) * 64).  passes: 3 = bf16x3, 1 = bf16.
// Sets *status = DESCO_ERANGE if a neighborhood does not fit a tile (results are then undefined).
int desco_internal_shmp_fused_layers(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                                     const uint8_t* edge_tri, int G, int pyg_batch_size, const float* feat, int input_dim,
                                     const float* w_pre, const void* w_layers_tc, int layers, int passes, float* emb_a,
                                     void* emb_img, float* pool, int emb_ld, void* workspace, int32_t* status,
                                     const int32_t* g_dev, cudaStream_t s);

// Y = act(X . W^T + b) (+ R) on tcgen05 (csrc/dense_tc.cu).  Wimg: pack_dense_tc images; act: 0 none, 1 relu, 2 leaky.
// Ximg != NULL: X comes as ready-made operand images (per block of 128 rows and 64-wide K atom: bf16 hi | mid | lo
// SWIZZLE_128B images, 3 x 16 KB) and is fetched by bulk async copy instead of being converted by the threads.
// m_dev / g_dev (may be NULL): device-resident row count of the stream-ordered form; M / G are then capacities.
int desco_internal_dense_tc(const float* X, const void* Ximg, int ldx, const void* Wimg, const float* bias, const float* R, int ldr, float* Y,
                            int ldy, int M, int K, int N, int nblk, int act, float slope, int passes, int32_t* status,
                            const int32_t* m_dev, cudaStream_t s);

// post_mp chain Z[G, K0] -> out[G, 64] in one launch, fp32 (csrc/readout.cu).  Weights row-major [in][out].
int desco_internal_readout_chain(const float* Z, int ldz, int K0, int G, const float* P0, const float* b0, const float* P1,
                                 const float* b1, const float* P2, const float* b2, const float* P3, const float* b3,
                                 float* out, const int32_t* g_dev, cudaStream_t s);

// Factorised count head for Q <= 32 queries in one launch, fp32 (csrc/readout.cu); DESCO_ERANGE for larger Q.
// Bq: [Q][256] scratch for the query half of the first Linear (q . W1b + b1, computed once per call by a small launch)
int desco_internal_count_head_fused(const float* emb_t, int G, const float* emb_q, int Q, const float* W1a, const float* W1b,
                                    const float* b1, const float* w2, const float* b2, float* pred, float* count,
                                    float* Bq, const int32_t* g_dev, cudaStream_t s);

/* bytes of one layer in the multi-tile weight blob (csrc/shmp_mt.cu): 3 K blocks x [hi | lo] images of a [64 n x 64 k]
 * block of Wc^T (tcpack.pack_b_operand) */
#define SHMP_MT_LAYER_BYTES (3 * 2 * 64 * 128)

int64_t desco_internal_shmp_mt_workspace_bytes(int num_rows, int num_neighborhoods);
int desco_internal_shmp_mt_layers(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                                  const uint8_t* edge_tri, int G, int V, int hetero, const int32_t* row_nbh,
                                  const int32_t* crow, const uint8_t* canon_code, const int32_t* quirk_row, float* hA,
                                  float* hB, float* emb_a, float* pool, float* cvec, int emb_ld, const float* w_layers,
                                  int64_t layer_floats, const void* w_layers_mt, int layers, int passes, void* workspace,
                                  int32_t* status, int anchored, cudaStream_t s);
// emb_a[g][layer*64 ..] = h[last row of neighborhood g]  (shmp.cu; the homogeneous model's centre rows)
void desco_internal_shmp_copy_last_rows(const int32_t* nbh_ptr, int G, const float* h, int layer, float* emb_a, int emb_ld,
                                        cudaStream_t s);
// cvec[g] = h_canonical^l[g] . Cw^l (shmp.cu)
void desco_internal_shmp_cvec(const float* emb_a, int emb_ld, int layer, const float* Cw, int G, float* cvec, cudaStream_t s);
