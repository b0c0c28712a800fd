/* desco_b200 - C ABI of the B200-native DeSCo inference hot path
 * (canonical neighborhood partition -> SHMP neighborhood counting -> gossip propagation).
 *
 * The reference (fuvty/DeSCo @ 4508f7a) is 100 % Python and has no FFI of its own; every entry point below cites the
 * reference function(s) it replaces (paths relative to the reference root).  All pointers are DEVICE pointers unless
 * the name says host; sizes are element counts; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Every function returns 0 on success or a negative errno-style code (DESCO_E*); nothing throws across the boundary.
 * Kernels that detect a violated precondition on the device write a DESCO_E* code into `status` (device int32,
 * caller-zeroed) which the caller reads at its next synchronisation point.
 *
 * Index types: int32 (target graphs up to 2^31-1 directed edges).  Features / counts: float32.
 */
#ifndef DESCO_B200_H
#define DESCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DESCO_MODE_HETERO 0    /* get_neigh_hetero    subgraph_counting/data.py:375-396 (default, hetero_graph=True)  */
#define DESCO_MODE_CANONICAL 1 /* get_neigh_canonical subgraph_counting/data.py:353-372 (hetero_graph=False)          */

#define DESCO_MODE_KHOP 2      /* k_neigh             subgraph_counting/data.py:329-338 (plain k-hop ball, no filter)      */

#define DESCO_PRECISION_FP32 0 /* fp32 FFMA everywhere: the 1e-4 parity path                                           */
#define DESCO_PRECISION_TF32X3 1 /* tcgen05 kind::tf32, 3-pass hi/lo split: fp32-level error on the tensor pipe        */
#define DESCO_PRECISION_BF16 2 /* tcgen05 kind::f16 (bf16 operands, fp32 accumulate): the 1e-2 variant                */

/* Library / build identification: returns e.g. "desco_b200 0.1 sm_100a". */
const char* desco_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Canonical partition + SHMP edge typing
 * Replaces: k_neigh / k_neigh_canonical / get_neigh_canonical / get_neigh_hetero  (data.py:329-396),
 *           the NeighborhoodDataset.process loop + indicator/index bookkeeping      (workload.py:243-260, 293-294),
 *           NetworkxToHetero                                                        (transforms.py:319-412),
 *           ToTconvHetero                                                           (transforms.py:180-255).
 *
 * Target graphs: one block-diagonal CSR. rowptr[N+1], col[M] (sorted rows, symmetric, no self loops),
 * graph_ptr[num_graphs+1] node ranges.  centres[num_centres]: dataset-global node ids, any order (the reference
 * iterates graph by graph, node ascending).  max_graph_nodes: caller's upper bound on the node count of any graph a
 * centre lives in (selects the warp-per-centre or CTA-per-centre kernel and sizes the shared-memory bitsets).
 *
 * Three calls: count -> scan -> (caller reads totals, allocates) -> fill.
 * ---------------------------------------------------------------------------------------------------------------- */

/* Pass 1: per centre, |V| and directed |E| of its canonical neighborhood.  out_ne == 0 marks a dropped neighborhood
 * (workload.py:253-256) and then out_nv is forced to 0.  out_centre_graph[i] = graph id of centre i. */
int desco_partition_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                          const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                          int32_t max_graph_nodes, int32_t* out_nv, int32_t* out_ne, int32_t* out_centre_graph,
                          int32_t* status, void* stream);

int64_t desco_partition_scan_workspace_bytes(int32_t num_centres);

/* Pass 2: exclusive scans.  keep_rank/node_off/edge_off [num_centres]; nbh_ptr[num_centres+1] (first G+1 used);
 * centre_out[num_centres] (first G used); indicator[num_centres] (uint8, may be NULL) == nx_neighs_indicator;
 * totals[3] = {G kept neighborhoods, V rows, E directed edges}. */
int desco_partition_scan(const int32_t* centres, const int32_t* nv, const int32_t* ne, int32_t num_centres,
                         int32_t* keep_rank, int32_t* node_off, int32_t* edge_off, int32_t* nbh_ptr,
                         int32_t* centre_out, uint8_t* indicator, int32_t* totals, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* Pass 3: emit the packed batch.  node_gid[V] ascending inside a neighborhood (canonical node = last row),
 * edge_ptr[V+1] CSR over batch rows, edge_col[E] batch-global row of the other endpoint (ascending inside a row),
 * edge_tri[E] 1 = "union_triangle", 0 = "union_tride". */
int desco_partition_fill(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                         const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                         int32_t max_graph_nodes, const int32_t* nv, const int32_t* ne, const int32_t* centre_graph,
                         const int32_t* node_off, const int32_t* edge_off, int32_t* node_gid, int32_t* edge_ptr,
                         int32_t* edge_col, uint8_t* edge_tri, int32_t* status, void* stream);

/* SHMP typing of an already-built batch / query set (ToTconvHetero applied to existing graphs, transforms.py:180-255;
 * also used for the query graphs, lightning_model.py:84-85).  Rows must have ascending edge_col. */
int desco_shmp_edge_types(const int32_t* edge_ptr, const int32_t* edge_col, int32_t num_rows, uint8_t* edge_tri,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DESCO_B200_H */
