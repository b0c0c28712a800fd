/* desco_b200 - C ABI of the B200-native DeSCo inference hot path
 * (canonical neighborhood partition -> SHMP neighborhood counting -> gossip propagation).
 *
 * The reference (fuvty/DeSCo @ 4508f7a) is 100 % Python and has no FFI of its own; every entry point below cites the
 * reference function(s) it replaces (paths relative to the reference root).  All pointers are DEVICE pointers unless
 * the name says host; sizes are element counts; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Every function returns 0 on success or a negative errno-style code (DESCO_E*: -22 EINVAL, -12 ENOMEM, -5 ECUDA,
 * -34 ERANGE, -105 ENOBUFS); nothing throws across the boundary.
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
#define DESCO_MODE_KHOP 2      /* k_neigh             subgraph_counting/data.py:329-338 (plain k-hop ball, no filter) */

#define DESCO_PRECISION_FP32 0   /* fp32 FFMA everywhere (layer-by-layer kernels)                                      */
#define DESCO_PRECISION_BF16X3 1 /* tcgen05 kind::f16, 3-pass bf16 hi/lo operand split, fp32 accumulate in TMEM:        */
                                 /* ~3e-6 from the fp32 oracle - the default 1e-4 parity path                           */
#define DESCO_PRECISION_BF16 2   /* tcgen05 kind::f16, single bf16 pass, fp32 accumulate: the 1e-2 variant              */

/* Library / build identification: returns e.g. "desco_b200 0.1 sm_100a". */
const char* desco_version(void);

/* Launch accounting and live device timing (no reference counterpart; measurement support for bench.py).
 * desco_kernel_launches: number of desco_b200 kernels launched by this process so far.
 * desco_profile_enable(1) starts recording one CUDA-event pair (on the launching stream) around every kernel of the
 * slots below; desco_profile_read synchronises those events and returns, per slot, the summed device milliseconds and
 * the number of launches, then clears the record. */
#define DESCO_PROF_PARTITION 0   /* partition count / fill kernels                 */
#define DESCO_PROF_SHMP_LAYER 1  /* shmp_layer_kernel (one launch per SHMP layer)  */
#define DESCO_PROF_SHMP_OTHER 2  /* plan, pre, cvec, pool, readout MLPs, count head */
#define DESCO_PROF_GOSSIP_L0 3   /* gossip layer-0 scalar sweep                    */
#define DESCO_PROF_GOSSIP_L1 4   /* gossip layer 1: gated-sweep gather kernel (FFMA path: the whole tile kernel) */
#define DESCO_PROF_GOSSIP_CHAIN 5 /* gossip layer 1 + post_mp: tcgen05 GEMM-chain kernel */
#define DESCO_PROF_SLOTS 6
int64_t desco_kernel_launches(void);
int desco_profile_enable(int32_t on);
int desco_profile_read(double* ms, int64_t* launches);

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
 * totals[3] = {G kept neighborhoods, V rows, E directed edges} - int32 scans: a caller that may pass more than 2^31
 * rows or edges worth of centres must check sum(nv), sum(ne) in 64 bits first (desco_b200.data.partition_batch does). */
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

/* Passes 1-3 in ONE call (what NeighborhoodDataset.process does per dataset, workload.py:215-294): count, scans, one
 * stream synchronisation to read the output sizes, fill.  The caller owns every buffer: `workspace`
 * (desco_partition_batch_workspace_bytes), the per-centre outputs (nbh_ptr[num_centres+1], centre_out / indicator /
 * centre_graph [num_centres]) and the packed batch at a capacity of its choice (node_gid[cap_rows],
 * edge_ptr[cap_rows+1], edge_col / edge_tri [cap_edges]).  totals_host[4] (host memory) receives {G, V, E, rows of the
 * largest neighborhood}.  Returns DESCO_ENOBUFS, with totals_host filled and nothing emitted, when V > cap_rows or
 * E > cap_edges: re-allocate and call again.  Returns DESCO_ERANGE (totals_host[1] = totals_host[2] = -1) when the exact
 * 64-bit row or edge total of the centre list does not fit the batch's int32 offsets: split the centre list. */
int64_t desco_partition_batch_workspace_bytes(int32_t num_centres);
int desco_partition_batch(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                          const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                          int32_t max_graph_nodes, void* workspace, int64_t workspace_bytes, int32_t* nbh_ptr,
                          int32_t* centre_out, uint8_t* indicator, int32_t* centre_graph, int32_t* node_gid,
                          int32_t* edge_ptr, int64_t cap_rows, int32_t* edge_col, uint8_t* edge_tri, int64_t cap_edges,
                          int32_t* totals_host, void* stream);

/* Stream-ordered form of desco_partition_batch (no host synchronisation, CUDA-graph capturable): count, scans, a capacity
 * guard and fill are enqueued back to back; the batch's own sizes stay ON THE DEVICE in sizes_dev (int32[16], caller
 * allocated): [12..15] = effective {G, V, E, rows of the largest neighborhood} read by the *_dev entry points below,
 * [0..3] the raw totals, [4] a status word, [8..11] the exact 64-bit row / edge sums.  When the totals exceed cap_rows /
 * cap_edges nothing is emitted, the effective sizes are 0 and [4] = DESCO_ENOBUFS (DESCO_ERANGE past 2^31): the caller
 * notices when it next reads sizes_dev and repeats with desco_partition_batch.  Graphs up to 409 600 nodes. */
int desco_partition_batch_async(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                                const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                                int32_t max_graph_nodes, void* workspace, int64_t workspace_bytes, int32_t* nbh_ptr,
                                int32_t* centre_out, uint8_t* indicator, int32_t* centre_graph, int32_t* node_gid,
                                int32_t* edge_ptr, int64_t cap_rows, int32_t* edge_col, uint8_t* edge_tri, int64_t cap_edges,
                                int32_t* sizes_dev, void* stream);

/* Large-graph variants of passes 1 and 3 (config 5: a 10M-node / 200M-directed-edge target, whose node bitsets no
 * longer fit shared memory; desco_partition_count returns DESCO_ERANGE there).  Same outputs, same reference
 * semantics (data.py:329-396).  An ordinary centre is served by one CTA with a hash set + member list + reached list
 * in shared memory; a centre whose ball overflows them by a TEAM of 16 co-resident CTAs (cooperative launch) with
 * member / reached bitmaps over the target graph in a per-team slice of `workspace`.  The SAME workspace
 * (desco_partition_large_workspace_bytes) must be passed to count and to fill, untouched in between: it carries the
 * tier of every centre, the list of team centres and the count pass's reached lists (fill skips the BFS for them).
 * fill also runs the SHMP typing of the whole batch (num_rows = totals[1]). */
int64_t desco_partition_large_workspace_bytes(int32_t max_graph_nodes, int32_t num_centres);
int desco_partition_large_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                                const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                                int32_t max_graph_nodes, int32_t* out_nv, int32_t* out_ne, int32_t* out_centre_graph,
                                int32_t* status, void* workspace, int64_t workspace_bytes, void* stream);
int desco_partition_large_fill(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                               const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                               int32_t max_graph_nodes, const int32_t* nv, const int32_t* ne, const int32_t* centre_graph,
                               const int32_t* node_off, const int32_t* edge_off, int32_t num_rows, int32_t* node_gid,
                               int32_t* edge_ptr, int32_t* edge_col, uint8_t* edge_tri, int32_t* status, void* workspace,
                               int64_t workspace_bytes, void* stream);
/* Capacities of the shared-memory tier (hash slots as log2, member-list entries, reached-list entries = power of two;
 * defaults 13 / 5120 / 4096); the last three arguments are ignored (the team tier has no capacity limit).  The tests
 * shrink the capacities to force the team tier on small graphs. */
int desco_partition_large_set_caps(int32_t log2_slots0, int32_t members0, int32_t reached0, int32_t log2_slots1,
                                   int32_t members1, int32_t reached1);
/* Profiling aid: clock64 cycles thread 0 of every shared-memory-tier CTA spent per phase since the last reset
 * (out[6]: frontier expansion, component, sort, induced degrees, edge emission, table wipe). */
int desco_partition_large_phase_cycles(uint64_t* out, int32_t reset);

/* SHMP typing of an already-built batch / query set (ToTconvHetero applied to existing graphs, transforms.py:180-255;
 * also used for the query graphs, lightning_model.py:84-85).  Rows must have ascending edge_col. */
int desco_shmp_edge_types(const int32_t* edge_ptr, const int32_t* edge_col, int32_t num_rows, uint8_t* edge_tri,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * SHMP neighborhood counting (forward)
 * Replaces: SAGEConv (gnn_model.py:362-404), BaseGNNCore.forward as expanded by to_hetero_old
 *           (gnn_model.py:230-277, lightning_model.py:371-421), BaseGNN.forward (gnn_model.py:58-109).
 *
 * Input: a packed batch (see desco_partition_fill).  hetero = 1: target neighborhoods (node types count/canonical,
 * canonical = last row of each neighborhood, 6 relations); hetero = 0: query graphs (one node type, 2 relations, no
 * anchor_mlp); hetero = 2: the homogeneous model of hetero_graph = False (workload.py:238-241, gnn_model.py:74-83): one
 * node type, the last row of every neighborhood is the centre (node_feature = 1 in `feat`) whose embedding goes through
 * anchor_mlp before the all-rows pooling (fp32 and multi-tile paths).  pyg_batch_size: size of the collated PyG batches the reference would have formed (config.py:255,
 * default 512; 0 = the whole input is one batch; < 0 = do not reproduce the quirk) - needed only to reproduce
 * SAGEConv's remove_self_loops on the bipartite relations (gnn_model.py:389-390), see DESIGN.md "reference quirks":
 * the dropped edge joins the canonical node and the FIRST count node of the neighborhood, which here is the lowest
 * node id; in the reference it is the first count node in networkx iteration order, which is the lowest id in most but
 * not all neighborhoods (225 of 232 in tests/golden/shmp_pipeline_ref.npz) and is not a function of the graph alone.
 * feat: [num_rows, input_dim] node features or NULL for ZeroNodeFeat (workload.py:431-440).
 *
 * Weight blobs (float32, K-major = transposed nn.Linear weights; built by desco_b200.gnn_model.pack_*):
 *   w_pre     per node type t in (count, canonical) [hetero] or (union_node): Wpre_t[input_dim][64], bpre_t[64]
 *   w_layers  per layer, desco_shmp_layer_weight_floats() floats:
 *               Wc[192][64]   = [ (U_m W_cc_tri)^T ; (U_m W_cc_tride)^T ; U_h^T ]          count destinations
 *               bias_c[64]    = U_m (sum of the biases of the 4 relations into count) + u
 *               Cw[64][128]   = [ (U_m W_ac_tri)^T | (U_m W_ac_tride)^T ]                  canonical -> count
 *               Wa[192][64], bias_a[64]                                                     canonical destinations
 *             (U = updates[l] weight split as [U_m | U_h], gnn_model.py:264; single-type graphs use Wc/bias_c only)
 *   w_readout Wanc[576][576], banc[576], P0[576][64], b0[64], P1[64][64], b1[64], P2[64][256], b2[256],
 *             P3[256][64], b3[64]                                    (anchor_mlp + post_mp, gnn_model.py:40-53)
 *   w_layers_tc  tensor-core form of w_layers for precision BF16X3 / BF16 (hetero batches only), per layer
 *             desco_shmp_tc_layer_bytes() bytes: the [tri | tride | self] weights as a 192 x 64 K-major operand, split in
 *             bf16 hi / lo and pre-swizzled (SWIZZLE_128B shared-memory images, desco_b200.tcpack), then fp32
 *             bias_c[64], bias_a[64], then Wa [192][64] and Cw [64][128] as warp-level mma.sync B fragments (bf16 hi / lo,
 *             desco_b200.tcpack.pack_mma_b_frags; same byte count as fp32).  May be NULL for precision FP32 (then w_layers is
 *             required); w_layers may be NULL for the tensor-core precisions.
 *   w_readout_tc  tensor-core form of the readout weights (biases still come from w_readout): operand images of Wanc
 *             (4 column blocks of 144), P0, P1, P2 (2 blocks of 128), P3 in that order, each
 *             [column block][64-wide K atom][hi | mid | lo] (3-way bf16 split, 6 tensor-core passes = fp32-grade
 *             products) as built by desco_b200.tcpack.pack_dense_tc.
 * out_emb: [num_neighborhoods, 64].  precision: DESCO_PRECISION_*.
 * status (device int32, caller-zeroed; required for the tensor-core precisions): DESCO_ERANGE when a neighborhood has
 * more than 128 rows - the fused kernel keeps a whole neighborhood in one 128-row tile - and the caller must rerun
 * with DESCO_PRECISION_FP32 (the layer-by-layer path has no size limit).
 * ---------------------------------------------------------------------------------------------------------------- */
int64_t desco_shmp_workspace_bytes(int32_t num_rows, int32_t num_neighborhoods, int32_t layers);
int64_t desco_shmp_layer_weight_floats(void);
int64_t desco_shmp_tc_layer_bytes(void);
int desco_shmp_forward(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                       int32_t num_neighborhoods, int32_t num_rows, int32_t hetero, int32_t pyg_batch_size,
                       const float* feat, int32_t input_dim, const float* w_pre, const float* w_layers,
                       const void* w_layers_tc, const float* w_readout, const void* w_readout_tc, int32_t layers,
                       int32_t hidden, float* out_emb, void* workspace, int64_t workspace_bytes, int32_t precision,
                       int32_t* status, void* stream);

/* Stream-ordered form of desco_shmp_forward for the fused tensor-core path (hetero batches, neighborhoods <= 128 rows):
 * cap_neighborhoods / cap_rows are CAPACITIES (they size the launch grids and the workspace), the batch's own
 * {G, V, E, max rows} are read on the device from sizes_dev (int32[4], e.g. desco_partition_batch_async's block + 12).
 * out_emb rows >= G are not written.  Together with desco_partition_batch_async and desco_count_head_dev the whole
 * partition -> SHMP -> count-head step is free of host round trips and can be captured in one CUDA graph. */
int desco_shmp_forward_dev(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                           int32_t cap_neighborhoods, int32_t cap_rows, const int32_t* sizes_dev, int32_t pyg_batch_size,
                           const float* feat, int32_t input_dim, const float* w_pre, const void* w_layers_tc,
                           const float* w_readout, const void* w_readout_tc, int32_t layers, int32_t hidden, float* out_emb,
                           void* workspace, int64_t workspace_bytes, int32_t precision, int32_t* status, void* stream);

/* The same forward for neighborhoods of ANY size on the tensor cores (csrc/shmp_mt.cu): features stay in HBM between
 * layers, a tile is 128 consecutive count rows whatever neighborhoods they belong to, the edge-type-split gather writes
 * the [128 x 192] operand into shared memory as bf16 hi/lo images and the product runs on tcgen05.  Serves Syn_1827-shaped
 * batches (neighborhoods of several hundred rows), the depth-2 balls of a power-law target (10^3 - 10^6 rows) and query
 * graphs (hetero = 0).  w_layers (fp32 blob, required: biases, Cw, Wa are read from it) as above; w_layers_mt: per layer
 * desco_shmp_mt_layer_bytes() bytes = for each 64-wide K block (tri | tride | self) the [64 n][64 k] block of Wc^T as a
 * bf16 hi / lo pre-swizzled operand image (desco_b200.tcpack.pack_b_operand).  precision: BF16X3 or BF16.
 * w_readout_tc is required for hetero batches. */
int64_t desco_shmp_mt_layer_bytes(void);
int desco_shmp_forward_mt(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                          int32_t num_neighborhoods, int32_t num_rows, int32_t hetero, int32_t pyg_batch_size,
                          const float* feat, int32_t input_dim, const float* w_pre, const float* w_layers,
                          const void* w_layers_mt, const float* w_readout, const void* w_readout_tc, int32_t layers,
                          int32_t hidden, float* out_emb, void* workspace, int64_t workspace_bytes, int32_t precision,
                          int32_t* status, void* stream);

/* Query-conditioned count head.  Replaces embed_to_count / the per-query loop of graph_to_count
 * (lightning_model.py:176-222, count_model :127-131):  pred[g,q] = count_model(cat(emb_target[g], emb_query[q])),
 * count = 2^pred - 1.  w_head: W1a[64][256] (target half of Linear(128,256)), W1b[64][256] (query half), b1[256],
 * w2[256], b2[1].  out_pred / out_count: [G, Q], either may be NULL. */
int64_t desco_count_head_workspace_bytes(int32_t num_neighborhoods, int32_t num_queries);
int desco_count_head(const float* emb_target, int32_t num_neighborhoods, const float* emb_query, int32_t num_queries,
                     const float* w_head, const void* w_head_tc, int32_t hidden, float* out_pred, float* out_count,
                     void* workspace, int64_t workspace_bytes, int32_t precision, int32_t* status, void* stream);

/* Stream-ordered form of desco_count_head (num_queries <= 32): cap_neighborhoods is a capacity, G = sizes_dev[0]. */
int desco_count_head_dev(const float* emb_target, int32_t cap_neighborhoods, const int32_t* sizes_dev, const float* emb_query,
                         int32_t num_queries, const float* w_head, int32_t hidden, float* out_pred, float* out_count,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Gossip propagation (forward)
 * Replaces: GossipConv (gnn_model.py:280-359), the GOSSIP branch of BaseGNNCore.forward (gnn_model.py:231-260),
 *           post_mp per node (gnn_model.py:102-103) and the per-query loop of GossipCountingModel.graph_to_count
 *           (lightning_model.py:613-628).  2 GossipConv layers, hidden 64, input_dim 1 (config.py:312-322).
 *
 * rowptr/col: the target CSR (symmetric, sorted, no self loops; node ids are batch-global so "j < i" is the
 * reference's edge_index[0] < edge_index[1], gnn_model.py:248).  x: [N, Q] neighborhood counts (row-major),
 * query_emb: [Q, 64].  out: [N, Q] = x + gossip correction.  out_gates: [2, Q] or NULL (_gate_value,
 * lightning_model.py:640-649).
 * Weight blobs are built by desco_b200.gnn_model.pack_gossip_weights (layout: csrc/gossip.cu WG_* / WQ_* offsets).
 *
 * The three stages are also exported separately so a node-range shard [node_begin, node_end) can run layer 1 after
 * the counts of its halo have been exchanged:  prepare_queries -> layer0 (any node range; needs x of the
 * neighbours) -> layer1 (needs s4 of the neighbours).  qvec: [Q, 256] floats, s4: [N, Q, 4] floats.
 * ---------------------------------------------------------------------------------------------------------------- */
/* Profiling aid: clock64 cycles thread 0 of every CTA of the tensor-core chain kernel (gossip_chain_kernel) spent per phase since the last
 * reset (out[6]: operand wait, x2, y1, y2, y4 GEMM + epilogue, spare). */
int desco_gossip_tc_phase_cycles(uint64_t* out, int32_t reset);
int64_t desco_gossip_weight_floats(void);
int64_t desco_gossip_query_weight_floats(void);
int64_t desco_gossip_workspace_bytes(int32_t num_nodes, int32_t num_queries);
int desco_gossip_prepare_queries(const float* query_emb, int32_t num_queries, const float* w_gossip_query, float* qvec,
                                 float* out_gates, void* stream);
int desco_gossip_layer0(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* x,
                        int32_t num_queries, const float* qvec, float* s4, void* stream);
int desco_gossip_layer1(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* s4,
                        int32_t num_queries, const float* qvec, const float* w_gossip, float* out, int32_t precision,
                        void* workspace, int64_t workspace_bytes, void* stream);
/* Query-grouped forms for the node-range-sharded forward (desco_b200/distributed.py; the reference has no multi-GPU
 * gossip, main.py:353-356).  The queries are cut into groups of `group_size` consecutive queries and s4 is laid out
 * group by group: group g is a dense [s4_rows][qc_g][4] block (qc_g = min(group_size, Q - g*group_size)) at float
 * offset 4*g*group_size*s4_rows, so that one group of one node range is a contiguous send buffer and the gathered group
 * a contiguous [s4_rows][qc_g][4] block: the halo all-gather of group g+1 runs under layer 1 of group g.
 * layer0_grouped writes rows [node_begin, node_end) of every group; layer1_group consumes ONE gathered group
 * (s4_group = that block, queries [query_begin, query_begin + group_queries)) and writes
 * out[i * out_stride + (q - query_begin)].  group_size >= Q and out_stride = Q give the plain layouts above. */
int desco_gossip_layer0_grouped(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end,
                                const float* x, int32_t num_queries, const float* qvec, float* s4, int32_t group_size,
                                int64_t s4_rows, void* stream);
int desco_gossip_layer1_group(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end,
                              const float* s4_group, int32_t query_begin, int32_t group_queries, const float* qvec,
                              const float* w_gossip, float* out, int32_t out_stride, int32_t precision, void* workspace,
                              int64_t workspace_bytes, void* stream);
/* Staging bytes desco_gossip_layer1 wants for a range of num_nodes nodes (0 for DESCO_PRECISION_FP32; for
 * DESCO_PRECISION_BF16X3 a 256-byte head + the operand images of up to 8192 tiles of 128 nodes x 1 query, 65 KB each -
 * any whole number of tiles works, the range is walked in chunks). */
int64_t desco_gossip_layer1_workspace_bytes(int32_t num_nodes, int32_t num_queries, int32_t precision);
int desco_gossip_forward(const int32_t* rowptr, const int32_t* col, int32_t num_nodes, const float* x,
                         int32_t num_queries, const float* query_emb, const float* w_gossip,
                         const float* w_gossip_query, float* out, float* out_gates, void* workspace,
                         int64_t workspace_bytes, int32_t precision, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Ground-truth canonical counts (labels)
 * Replaces: MatchSubgraphWorker (workload.py:327-348, networkx VF2 per (target, query), every mapping credited to
 *           max(vmap.keys())), Workload.compute_groundtruth (workload.py:551-726) and the division by SymmetricFactor
 *           (data.py:61-88): out_counts[node, q] = number of node sets whose induced subgraph is isomorphic to query q and
 *           whose largest node is `node` (int64, caller-zeroed).  Connected queries of 3..5 nodes.
 * lut3 / lut4 / lut5: uint8[8] / [64] / [1024], adjacency pattern of a node tuple (bit of pair (i, j), i < j, at position
 * j(j-1)/2 + i) -> query column, 255 = none (desco_b200.groundtruth.pattern_tables).  Graphs up to 1024 nodes (the
 * adjacency bit-matrix of a graph lives in shared memory); status receives DESCO_ERANGE otherwise.
 * ---------------------------------------------------------------------------------------------------------------- */
int desco_groundtruth_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                            int32_t max_graph_nodes, const uint8_t* lut3, const uint8_t* lut4, const uint8_t* lut5,
                            int32_t num_queries, int32_t max_query_nodes, int64_t* out_counts, int32_t* status,
                            void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Stand-alone message-passing primitives behind the callable leaf modules (csrc/conv.cu).  The hot path runs whole
 * layer stacks fused; these serve code written against the reference's module API:
 *   SAGEConv.forward   gnn_model.py:372-404  = desco_spmm_sum (propagate, aggr = "add") + desco_train_dense (lin)
 *   GossipConv.forward gnn_model.py:303-359  = desco_gossip_gate, desco_train_dense (lin_com per node), desco_spmm_sum with
 *                                              the per-edge gate weights, desco_train_dense (lin_update)
 * desco_spmm_sum: out[i, 0:width] = sum over CSR row i (rowptr[n_dst+1], col) of edge_w[e] * x[col[e], 0:width]
 * (edge_w NULL = 1).  desco_gossip_gate: gate[q] = LeakyReLU(sigmoid(w2 . sigmoid(W1 qemb[q] + b1) + b2)),
 * W1 [hidden][emb_channels] row-major (nn.Linear layout), GossipConv.lin_gate / _gate_value (gnn_model.py:294-301,353).
 * ---------------------------------------------------------------------------------------------------------------- */
int desco_spmm_sum(const int32_t* rowptr, const int32_t* col, const float* edge_w, int32_t n_dst, const float* x,
                   int32_t ldx, int32_t width, float* out, int32_t ldo, void* stream);
int desco_gossip_gate(const float* query_emb, int32_t num_queries, int32_t emb_channels, const float* w1, const float* b1,
                      int32_t hidden, const float* w2, const float* b2, float* gate, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step of the gossip model (csrc/gossip_train.cu)
 * Replaces: GossipCountingModel.train_forward / criterion (lightning_model.py:585-608, 630-635) and autograd through
 *           GossipConv (gnn_model.py:294-348).  The matrix work runs on desco_train_dense / desco_train_wgrad /
 *           desco_spmm_sum (whose adjoint on a symmetric edge set is the same call with the two gate weights swapped);
 *           these are the remaining pieces, all stream-ordered with the gate as a DEVICE scalar:
 *   gated_mix      out = g a + (1 - g) b                              (the two directional aggregates -> the message sum)
 *   gate_grad      *dgate += <d, a - b>                               (d loss / d gate)
 *   gate_backward  gradients of lin_gate (Linear, Sigmoid, Linear, Sigmoid, LeakyReLU) from *dgate, accumulated (+=)
 *   train_dropout  x *= mask ? scale : 0                              (F.dropout / nn.Dropout in training mode)
 *   gossip_loss    pred = c + out;  *loss += sum log2(|pred - y| + 1);  dout = d loss / d out   (strided columns)
 * ---------------------------------------------------------------------------------------------------------------- */
int desco_gossip_gated_mix(const float* a, const float* b, const float* gate, float* out, int64_t n, void* stream);
int desco_gossip_gate_grad(const float* d, const float* a, const float* b, int64_t n, float* dgate, void* stream);
int desco_gossip_gate_backward(const float* query_emb, int32_t emb_channels, const float* w1, const float* b1, int32_t hidden,
                               const float* w2, const float* b2, const float* dgate, float* dw1, float* db1, float* dw2,
                               float* db2, void* stream);
int desco_train_dropout(float* x, const uint8_t* mask, float scale, int64_t n, void* stream);
int desco_gossip_loss(const float* c, int32_t ldc, const float* out, int32_t ldo, const float* y, int32_t ldy, int32_t n,
                      float* pred, float* dout, int32_t ldd, float* loss, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training step of the neighborhood-counting model (BASELINE config 3)
 * Replaces: NeighborhoodCountingModel.train_forward (lightning_model.py:228-254), criterion (:285-289),
 *           configure_optimizers / torch.optim.Adam (:160-173) and the autograd graph through gnn_model.py:58-109,
 *           230-277, 362-404.  fp32.  The host side (desco_b200/training.py) sequences these primitives; parameters stay
 *           in their torch layout ([out][in] row-major) and gradients are accumulated (+=) straight into .grad buffers.
 * Node features live per node type in compact matrices: count rows [Vc, .] (all rows for single-type query graphs),
 * canonical rows [G, .]; compact count row of packed row r in neighborhood g is r - g.
 * ---------------------------------------------------------------------------------------------------------------- */

/* row_nbh[V]: packed row -> neighborhood; crow_nbh[Vc]: compact count row -> neighborhood; quirk_row[G]: the packed
 * row whose edge to the canonical node SAGEConv's remove_self_loops drops (gnn_model.py:389-390), or -1. */
int desco_train_plan(const int32_t* nbh_ptr, int32_t num_neighborhoods, int32_t hetero, int32_t pyg_batch_size,
                     int32_t* row_nbh, int32_t* crow_nbh, int32_t* quirk_row, void* stream);

/* transpose = 0: SAGEConv sum-aggregation (gnn_model.py:392), split by relation:
 *   ac[count row][slot*64..] = sum of x over its neighbours of (source type, SHMP type) = slot,
 *   slot = 2*(source is canonical) + (tride ? 1 : 0)  (4 slots hetero, 2 otherwise);  aa[g][slot*64..], slot = tride.
 * transpose = 1: the adjoint: xc / xa += gathered ac / aa (the edge set and the types are symmetric). */
int desco_train_aggregate(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                          const uint8_t* edge_tri, const int32_t* row_nbh, const int32_t* quirk_row, int32_t num_rows,
                          int32_t hetero, int32_t transpose, float* xc, int32_t ldc, float* xa, int32_t lda, float* ac,
                          int32_t ld_ac, float* aa, int32_t ld_aa, void* stream);

/* y[m, n] (+)= act( sum_b x[b][m, 0:64] . W_b^T + sum_i bias[i] ).  x, ldx, w, ldw, bias are HOST arrays of device
 * pointers / strides, one entry per 64-wide K block (<= 12) / bias (<= 4).  w_nmajor = 1: W_b[n*ldw + k] (a torch
 * Linear weight, forward);  0: W_b[k*ldw + n] (the same weight used for the data gradient).  act: 0 none, 1 relu,
 * 2 leaky(slope).  accumulate = 1 adds to y (act must be 0).  n % 64 == 0. */
int desco_train_dense(const float* const* x, const int32_t* ldx, const float* const* w, const int32_t* ldw,
                      int32_t num_blocks, int32_t w_nmajor, const float* const* bias, int32_t num_bias, float* y,
                      int32_t ldy, int32_t m, int32_t n, int32_t act, float slope, int32_t accumulate, void* stream);

/* dw[b][n*ldw[b] + k] += sum_m dy[m][n] x[m][64 b + k]  (b < num_blocks);  db[i][n] += sum_m dy[m][n]  (i < num_db). */
int desco_train_wgrad(const float* x, int32_t ldx, int32_t num_blocks, const float* dy, int32_t ldy, int32_t n, int32_t m,
                      float* const* dw, const int32_t* ldw, float* const* db, int32_t num_db, void* stream);

/* dx[m][c] *= act'(fwd[m][c]), derivative read off the activation output (act 1 relu, 2 leaky(slope)). */
int desco_train_act_backward(float* dx, int32_t ldd, const float* fwd, int32_t ldf, int32_t m, int32_t c, int32_t act,
                             float slope, void* stream);
/* y[m][0:c] = bias (pre_mp on ZeroNodeFeat inputs);  db[0:c] += column sums of dy. */
int desco_train_fill_rows(float* y, int32_t ldy, int32_t m, int32_t c, const float* bias, void* stream);
int desco_train_colsum(const float* dy, int32_t ldy, int32_t m, int32_t c, float* db, void* stream);

/* backward = 0: pooled[g] = sum of the count rows of g (+ z_a[g] if not NULL)   (global_add_pool, gnn_model.py:107);
 * backward = 1: emb_c[row] = pooled[neighborhood of row]  (pooled holds the gradient). */
int desco_train_pool(const int32_t* nbh_ptr, const int32_t* crow_nbh, int32_t num_neighborhoods, int32_t num_count_rows,
                     int32_t hetero, int32_t c, int32_t backward, float* emb_c, int32_t ldc, const float* z_a,
                     int32_t lda, float* pooled, int32_t ldp, void* stream);

/* Count head on T = t.W1[:, :64]^T [G,256] and Bq = q.W1[:, 64:]^T + b1 [Q,256] (Q <= 32):
 *   pred[g,q] = w2 . leaky_0.01(T[g] + Bq[q]) + b2;  with y != NULL also the training loss
 *   loss += mean_q mean_g smooth_l1(pred, log2(y + 1)) and dpred = d loss / d pred. */
int desco_train_head_loss(const float* t, const float* bq, const float* w2, const float* b2, const float* y,
                          int32_t num_neighborhoods, int32_t num_queries, float* pred, float* dpred, float* loss,
                          void* stream);
/* dt[G,256] = d loss / d T (written);  dbq[Q,256], dw2[256], db2[1] += their gradients. */
int desco_train_head_backward(const float* t, const float* bq, const float* w2, const float* dpred,
                              int32_t num_neighborhoods, int32_t num_queries, float* dt, float* dbq, float* dw2,
                              float* db2, void* stream);

/* torch.optim.Adam step (no amsgrad) over a flat fp32 buffer; step counts from 1. */
int desco_train_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t step, void* stream);

/* Phase profile of the fused SHMP layer kernel (measurement support, no reference counterpart): clock64 cycles summed
 * over all CTAs since the last reset, as seen by thread 0 of each CTA; counted only by the instantiation that is launched
 * when DESCO_FUSED_PHASE_TIMING=1 is in the environment (the production kernel carries no counters).  out[10] = {tile
 * setup, pool stage A up to its barrier, weight prefetch of warp 0, canonical rows (mma.sync), wait for the MMA,
 * TMEM -> shared memory, gather, pool stage B up to its barrier; then, as seen by the issuing thread: wait for the layer's
 * weight images, issue of the 12 tcgen05.mma}. */
int desco_shmp_fused_phase_cycles(uint64_t* out, int32_t reset);

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core self test (no reference counterpart): d[128][n] = a[128][64] . B[n][64]^T on tcgen05 with the bf16 hi/lo
 * split (passes = 1: hi.hi only, 3: hi.hi + lo.hi + hi.lo).  b_image is the pre-swizzled operand image built by
 * desco_b200.tcpack.pack_b_operand (n*128 bytes hi, then n*128 bytes lo).  n % 32 == 0, n <= 256.
 * ---------------------------------------------------------------------------------------------------------------- */
int desco_tc_selftest(const float* a, const void* b_image, int32_t n, int32_t passes, float* d, int32_t* status,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DESCO_B200_H */
