/* seggroup_b200 — C-ABI of the B200-native SegGroup hot path (libseggroup_b200.so).
 *
 * Every entry point takes plain DEVICE pointers (unless the name ends in `_host`), sizes and the
 * CUDA stream to launch on (`void* stream` = cudaStream_t; NULL = legacy default stream).  No entry
 * point allocates device memory: callers pass workspaces sized by the matching `*_ws_bytes`.
 * Return value: 0 = ok, negative = `sgb_status`.  Nothing here synchronises the stream unless the
 * comment says so; the work is complete when the stream reaches the point after the call.
 *
 * Index type is int32 everywhere on the device (the reference uses int64 torch tensors in
 * seggroup/model.py and int32 in the kpconv C++ cores; the Python layer converts at the boundary).
 * Each declaration cites the reference code it replaces (paths relative to the reference root).
 */
#ifndef SEGGROUP_B200_H
#define SEGGROUP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SGB_OK = 0,
    SGB_ERR_INVALID = -1,        /* bad argument (null pointer, negative size, unsupported width) */
    SGB_ERR_WORKSPACE = -2,      /* workspace too small */
    SGB_ERR_CUDA = -3,           /* a CUDA runtime call failed: see sgb_last_cuda_error() */
    SGB_ERR_UNSUPPORTED = -4,    /* shape outside what the kernels are built for */
    SGB_ERR_NO_DEVICE = -5
} sgb_status;

int sgb_version(void);                       /* 100*major + minor */
const char* sgb_status_string(int status);
int sgb_last_cuda_error(void);               /* cudaError_t of the last SGB_ERR_CUDA on this thread */
const char* sgb_last_cuda_error_string(void);
int sgb_device_arch(int* major, int* minor); /* compute capability of the current device */
long long sgb_launch_count(void);            /* kernels of this library launched by this process so far (all threads) */

/* ---------------------------------------------------------------------------------------------
 * Primitives (used by the ops below; exported for the parity tests)
 * ------------------------------------------------------------------------------------------- */
size_t sgb_scan_ws_bytes(int n);
/* out[i] = sum(in[0..i-1]); out has n+1 entries (out[n] = total).  in/out may not alias. */
int sgb_exclusive_scan_i32(const int* in, int* out, int n, void* ws, size_t ws_bytes, void* stream);

/* Dense fp32 GEMM on the tcgen05 tensor cores (TF32 x 3 split, fp32-level accuracy): C [M,N] = A [M,K] * B [N,K]^T,
 * the shape of nn.Linear (the GCN fc of seggroup/model.py:146-151).  N multiple of 16 and <= 256, K multiple of 4. */
int sgb_gemm_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, void* stream);
/* a12  the GCN layer's dense part relu(fc(.)) (seggroup/model.py:146-151; fc = nn.Linear(bias=False)): C = [relu](A B^T) with the
 * ReLU fused into the TMEM epilogue. */
int sgb_linear_tf32x3(const float* A, const float* B, float* C, int M, int N, int K, int relu, void* stream);
/* split-K form for the weight gradient dW = dZ^T X (K = number of clusters, thousands; M, N <= 256): Cpart [ceil(K / k_per_split), M, N],
 * one slab per K range of k_per_split (multiple of 32) columns; the caller sums the slabs in order. */
int sgb_gemm_tf32x3_splitk(const float* A, const float* B, float* Cpart, int M, int N, int K, int k_per_split, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a10  segment pooling: point features -> segment features
 * replaces seggroup/model.py:278-288 `aggregate_cluster_feature` (use_avg=False at every call
 * site: 770, 793, 815, 834, 856, 470, 507) and the per-instance max of model.py:912-919.
 *
 * Segment s owns rows members[offsets[s] .. offsets[s+1]) of feat (members == NULL: identity, i.e.
 * feat rows are already in segment-major order).  out[s,c] = max over the rows, argmax[s,c] = the
 * FIRST maximal row in member-list order (what `torch.max(dim=0)` over the gathered rows returns,
 * SURVEY.md 9.2 #14) as a row id of feat.  NaN propagates like torch.max.  Empty segments are
 * invalid input.  ws: sgb_segment_pool_ws_bytes(S, C).
 * ------------------------------------------------------------------------------------------- */
size_t sgb_segment_pool_ws_bytes(int S, int C);
int sgb_segment_pool_max_fwd(const float* feat, int n_rows, int C, const int* members, int n_members,
                             const int* offsets, int S, float* out, int* argmax,
                             void* ws, size_t ws_bytes, void* stream);
/* grad_feat[argmax[s,c], c] += grad_out[s,c]; grad_feat must be zero-filled (or hold a running sum).
 * Deterministic when the segments are disjoint (every row has at most one owner), as in the model. */
int sgb_segment_pool_max_bwd(const float* grad_out, const int* argmax, int S, int C,
                             float* grad_feat, void* stream);
/* use_avg variant (seggroup/model.py:282-284, `torch.mean(Feat_old[indexs], dim=0)`; never enabled by the reference's callers):
 * out[s,c] = mean over the rows of segment s (member-list order, fp32); an empty segment gives NaN as torch.mean does.
 * Backward: grad_feat[row, c] += grad_out[s,c] / n_s for every row of segment s; grad_feat must be zero-filled (or hold a sum). */
int sgb_segment_pool_mean_fwd(const float* feat, int n_rows, int C, const int* members, int n_members,
                              const int* offsets, int S, float* out, void* stream);
int sgb_segment_pool_mean_bwd(const float* grad_out, int S, int C, const int* members, int n_members,
                              const int* offsets, float* grad_feat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5/a7  kNN inside clusters
 * replaces seggroup/model.py:30-36 `knn` and 512-522 `get_knn` (k = 20).
 *
 * order[cl_off[c] .. cl_off[c+1]) are the point ids of cluster c in member order.  For a cluster
 * with n > k members: knn[p, :] = the k members with the largest score
 *   score(i,j) = (-|x_j|^2 - (-2 * <x_i,x_j>)) - |x_i|^2
 * evaluated exactly as torch-CPU does in fp32 (<.,.> = fma(z,z,fma(y,y,x*x)), |x|^2 = (x*x+y*y)+z*z),
 * ranked (score desc, member position asc).  For n <= k: the first n columns are all members in
 * member order, the remaining columns are 0 (the reference leaves them pointing at global point 0).
 * xyz rows have `stride` floats.  knn is [N,k] int32, every row is written.  Exact (identical to a brute force over the
 * cluster): clusters are sorted along their longest axis and every query sweeps outwards until the axis gap alone
 * rules out a better score (csrc/cluster_knn.cu).
 * ------------------------------------------------------------------------------------------- */
size_t sgb_cluster_knn_ws_bytes(int N, int S);
int sgb_cluster_knn(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                    int k, int* knn, void* ws, size_t ws_bytes, void* stream);
/* scene batch: neighbour ids are written RELATIVE to the scene's first point (scene_pt_off [n_scenes+1], device), i.e. exactly
 * the lists the reference builds for that scene alone, incl. the "unfilled columns point at point 0" rule of model.py:513-518. */
int sgb_cluster_knn_scenes(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                           int k, int* knn, const int* scene_pt_off, int n_scenes, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  per-cluster fixed-size clouds
 * replaces seggroup/model.py:398-426 `get_cluster_pointcloud`, 329-395 `farthest_point_sampling`.
 *
 * cloud_idx[c, 0..P) = point ids: the members tiled floor(P/n) times followed by P mod n farthest
 * point picks (start member 0, skip_initial, fp32 squared distances without FMA contraction,
 * argmax ties -> lowest member position, trailing picks equal to member 0 replaced by the leading
 * picks, model.py:407-412).  status[0] |= 1 if a cluster with P mod n > 0 has all members
 * coincident (the reference raises there).
 * sgb_cluster_cloud_transform writes clouds[c,p,0..6): xyz minus the mean over the P rows, divided
 * by max |xyz| of the centred cloud (model.py:421-423); rgb copied.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_cluster_cloud_ws_bytes(int N);
int sgb_cluster_cloud_indices(const float* xyz, int stride, int N, const int* order, const int* cl_off, int S,
                              int P, int* cloud_idx, int* status, void* ws, size_t ws_bytes, void* stream);
int sgb_cluster_cloud_transform(const float* data6, const int* cloud_idx, int S, int P, float* clouds,
                                void* stream);

/* a8  replaces seggroup/model.py:429-436 `combine_centralized_pointcloud`:
 * x9[p] = (data6[p], xyz[p] - mean xyz of p's cluster). */
size_t sgb_centralize_ws_bytes(int N, int S);
int sgb_centralize(const float* data6, int N, const int* order, const int* cl_off, int S, float* x9,
                   void* ws /* sgb_centralize_ws_bytes(N, S), 8-byte aligned: 64-bit fixed-point coordinate sums per cluster
                               + the cluster id of every point (pass 1 scatters it, pass 2 streams in point order) */,
                   size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5/a6  structural-layer MLP on the 64-point segment clouds
 * replaces seggroup/model.py:39-80 (`get_graph_feature1`, `MLP1.forward`, BatchNorm2d in training
 * mode = statistics of this scene over S*64*10 edge activations).
 * clouds [S,64,6] -> feat [S,128] = cat(max over points, mean over points) of max_k lrelu(BN(W e)).
 * W [64,6]; knn_idx [S,64,10] local neighbour ids (score desc, index asc); arg_pt [S,64] (may be NULL).
 * stats [4][64] = (batch mean, 1/sqrt(var+eps), gamma/sqrt(var+eps), beta); var [64] biased batch
 * variance (for the running-stat update); mom [27] fp64 input moments (kept for the backward).
 * ------------------------------------------------------------------------------------------- */
size_t sgb_mlp1_ws_bytes(int S);
int sgb_mlp1_fwd(const float* clouds, int S, const float* W, const float* gamma, const float* beta,
                 float* feat, int* knn_idx, int* arg_pt, float* stats, float* var, double* mom,
                 void* ws, size_t ws_bytes, void* stream);

/* Parameter gradients of sgb_mlp1_fwd (the clouds carry no gradient): g [S,128] -> gW [64,6], ggamma, gbeta [64].
 * Analytic training-mode BatchNorm backward from the stored input moments (see csrc/edgeconv_bwd.cu). */
size_t sgb_mlp1_bwd_ws_bytes(int S);
int sgb_mlp1_bwd(const float* g, const float* clouds, const int* knn_idx, const int* arg_pt, int S, const float* W,
                 const float* stats, const double* mom, float* gW, float* gg, float* gb,
                 void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9  EdgeConv point MLPs over the kNN graph
 * replaces seggroup/model.py:83-138 (`get_graph_feature2`, `MLP2.forward`, `MLP3.forward`).
 * x9 [N,9], knn [N,20] -> out [N,64] = max_k lrelu(BN1(W1 e)) (two_layer = 0, MLP2) or
 * max_k lrelu(BN2(W2 lrelu(BN1(W1 e)))) (two_layer = 1, MLP3), e = (x_j - x_i, x_i).
 * argk [N,64] uint8 (may be NULL): arg-max edge per point and channel, needed by the backward.
 * W1 [64,18], W2 [64,64]; stats1, stats2, var1, var2 as in sgb_mlp1_fwd; mom1 [189] / mom2 [4160] fp64 moments of the
 * layer inputs (NULL allowed when no backward follows); ctr_out [18] = centre e0 the first moments were taken about
 * (NULL allowed).  With two_layer = 1 and mom2 = NULL the 64x64 contraction runs on the tcgen05 tensor cores (TF32 x 3
 * split, fp32-level accuracy) in one fused pass that also yields the BatchNorm-2 statistics (csrc/edgeconv_tc.cu).
 * ------------------------------------------------------------------------------------------- */
size_t sgb_edgeconv_ws_bytes(int N, int two_layer);
int sgb_edgeconv_fwd(const float* x9, const int* knn, int N, int two_layer,
                     const float* W1, const float* gamma1, const float* beta1,
                     const float* W2, const float* gamma2, const float* beta2,
                     float* out, unsigned char* argk, float* stats1, float* var1, double* mom1,
                     float* stats2, float* var2, double* mom2, float* ctr_out,
                     void* ws, size_t ws_bytes, void* stream);
/* Parameter gradients of sgb_edgeconv_fwd followed by the point -> segment max pooling (model.py:793, 834):
 * g [S,64] = gradient of the pooled features, arg [S,64] = arg-max point ids from sgb_segment_pool_max_fwd,
 * argk [N,64] = arg-max edge per point and channel from the forward.  x9 carries no gradient. */
size_t sgb_edgeconv_bwd_ws_bytes(int N, int S, int two_layer);
int sgb_edgeconv_bwd(const float* g, const int* arg, const unsigned char* argk, int S, const float* x9, const int* knn, int N,
                     int two_layer, const float* W1, const float* stats1, const double* mom1, const float* e0,
                     const float* W2, const float* stats2, const double* mom2,
                     float* gW1, float* gg1, float* gb1, float* gW2, float* gg2, float* gb2,
                     void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a2  union-find state of the segment graph
 * replaces seggroup/model.py:169-214 `DisjointSet` and the graph initialisation of 712-721.
 * The state is SEGMENT level (a cluster is always a concatenation of whole level-1 segments,
 * model.py:191): uf = int[6][S1] = parent, next, tail, pnum, ins, sem; member lists are linked lists of
 * level-1 segments in the reference's list order (root = head).
 * seg_off [S1+1] / seg_members [N]: CSR of the input over-segmentation (seg.json), segments in ascending
 * root-point order; weak_label [N,2] int32 (sem, ins), -1 = unlabeled.
 * Outputs: seg_of_point [N], seg_of_pos [N] (segment of the q-th entry of seg_members), uf.
 * ------------------------------------------------------------------------------------------- */
int sgb_scene_init(const int* seg_off, const int* seg_members, const int* weak_label, int N, int S1,
                   int* seg_of_point, int* seg_of_pos, int* uf, void* stream);

/* Level construction: replaces `get_cluster_list` (model.py:209-214) and the cluster/cluster_map/
 * cluster_unmap dict loops (727-732, 759-768, 804-813, 845-854).  Clusters are enumerated by ascending
 * root point id.  All outputs are sized for S1 (offset arrays S1+1); counts[0] <- number of clusters.
 *   roots [S]: root level-1 segment of every cluster;  seg2cl [S1]: cluster of every level-1 segment;
 *   cl_seg_off/cl_seg_list: level-1 segments of every cluster in member-list order;
 *   cl_pt_off/order [N]: point ids of every cluster in member-list order;
 *   cl_ins, cl_sem, cl_rootpt [S]: weak labels and root point id of every cluster. */
size_t sgb_level_ws_bytes(int S1);
int sgb_level_build(const int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                    int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                    int* cl_ins, int* cl_sem, int* cl_rootpt, int* counts, void* ws, size_t ws_bytes, void* stream);
/* Scene batches.  seggroup/train.py:92 forces batch size 1 (one scene per rank per step, model.py:684-693 unpacks it); the
 * `*_scenes` entry points take SEVERAL scenes concatenated into one block-diagonal graph (point / segment ids offset by the
 * scene's base, no edge between scenes) so that one launch serves the whole batch, and keep every per-scene rule of the
 * reference: scene_seg_off / scene_cl_off / scene_pt_off [n_scenes+1] (device) are the level-1 segment, cluster and point
 * ranges of the scenes.  With n_scenes == 1 and NULL scene arrays they are the single-scene entry points. */
int sgb_level_build_scenes(const int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                           int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                           int* cl_ins, int* cl_sem, int* cl_rootpt, int* counts,
                           const int* scene_seg_off, int n_scenes, int* scene_cl_off, void* ws, size_t ws_bytes, void* stream);
/* cluster_new_to_old (model.py:760-768): old2new [n_old]; child_off [n_new+1] / child_list [n_old] = old dense
 * cluster ids of every new cluster, ascending. */
size_t sgb_children_ws_bytes(int n_new);
int sgb_level_children(const int* roots_old, int n_old, const int* seg2cl_new, int n_new, int* old2new,
                       int* child_off, int* child_list, void* ws, size_t ws_bytes, void* stream);

/* a3  replaces seggroup/model.py:291-302 `update_adj`: map both endpoints (map[old id] = new dense id
 * < S_new), drop self edges, order each pair, unique rows in lexicographic order.  counts[1] <- rows. */
size_t sgb_update_adj_ws_bytes(int S_new);
int sgb_update_adj(const int* edges, int E, const int* map, int S_new, int* adj_out, int* counts,
                   void* ws, size_t ws_bytes, void* stream);

/* symmetric CSR of a unique (u<v) edge list: row i = (neighbour, edge id) sorted by neighbour.
 * row_off [S+1], nbr/eid [2A].  Gives every gather-reduce below a fixed summation order. */
size_t sgb_sym_csr_ws_bytes(int S);
int sgb_sym_csr(const int* adj, int A, int S, int* row_off, int* nbr, int* eid, void* ws, size_t ws_bytes, void* stream);

/* One whole clustering level in one call (csrc/level_step.cu): [group_nearby (mode 0) | group_unlabeled_step (mode 1) | no
 * grouping (mode 2)] -> level_build -> level_children -> update_adj -> sym_csr, i.e. seggroup/model.py:752-770 / 802-815 /
 * 843-856 / 441-470 between "distances are known" and "features can be pooled".  Synchronises the stream twice (cluster
 * count, edge count).  edges [E,2] / map: the edge list re-mapped into adj_new (map == NULL: old2new of this step);
 * roots_old == NULL on the first level (no children CSR).  counts_host [4] (HOST memory) <- S_new, A_new, number of
 * unlabeled clusters, status word.  Cluster-sized outputs are sized for S1, adj_new for E rows, csr_nbr / csr_eid for 2 E. */
size_t sgb_level_step_ws_bytes(int S1, int S_old);
int sgb_level_step(int mode, const int* adj_old, int A_old, const int* roots_old, int S_old, const float* dist, float th,
                   int sweep_cap, const int* csr_off_old, const int* csr_nbr_old, const int* csr_eid_old,
                   const int* edges, int E, const int* map,
                   int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                   int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                   int* cl_ins, int* cl_sem, int* cl_rootpt, int* old2new, int* child_off, int* child_list,
                   int* adj_new, int* csr_off, int* csr_nbr, int* csr_eid,
                   int* status, int* counts_dev, int* counts_host, void* ws, size_t ws_bytes, void* stream);
/* scene batch: counts_dev / counts_host hold 4 + n_scenes + 1 ints (the four counters, then scene_cl_off_new); the order-dependent
 * replays run one CTA per scene (max_scene_segs / max_scene_cl_old size their shared-memory state). */
int sgb_level_step_scenes(int mode, const int* adj_old, int A_old, const int* roots_old, int S_old, const float* dist, float th,
                          int sweep_cap, const int* csr_off_old, const int* csr_nbr_old, const int* csr_eid_old,
                          const int* edges, int E, const int* map,
                          int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                          int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                          int* cl_ins, int* cl_sem, int* cl_rootpt, int* old2new, int* child_off, int* child_list,
                          int* adj_new, int* csr_off, int* csr_nbr, int* csr_eid,
                          int* status, int* counts_dev, int* counts_host,
                          const int* scene_seg_off, const int* scene_cl_off_old, int* scene_cl_off_new, int n_scenes,
                          int max_scene_segs, int max_scene_cl_old, void* ws, size_t ws_bytes, void* stream);

/* a11  replaces seggroup/model.py:269-274 `calculate_distance` (F.pairwise_distance: ||a - b + 1e-6||_2). */
int sgb_edge_dist_fwd(const float* feat, int C, const int* adj, int A, float* dist, void* stream);
int sgb_edge_dist_bwd(const float* feat, int S, int C, const int* adj, int A, const float* dist, const float* gdist,
                      const int* row_off, const int* eid, float* gfeat /* += */, void* stream);

/* a12  replaces seggroup/model.py:305-309 `build_similarity_matrix` + the normalise/aggregate half of
 * `GCN.forward` (146-149): AX = ((I + sym(sims)) / rowsum) X as a CSR gather-reduce; the dense
 * fc + relu that follows is a plain library GEMM on the caller's side. */
int sgb_gcn_agg_fwd(const float* X, int S, int C, const float* sims, const int* row_off, const int* nbr, const int* eid,
                    float* AX, float* rowsum, void* stream);
int sgb_gcn_agg_bwd(const float* dAX, const float* X, const float* AX, int S, int C, const float* sims, const float* rowsum,
                    const int* adj, int A, const int* row_off, const int* nbr, const int* eid,
                    float* dX, float* dsims, void* stream);

/* a13  replaces seggroup/model.py:218-258 `group_nearby_clusters`: edge-order replay of the unions
 * (skip iff dist > th), then the small-cluster (< 5 points) sweeps.  adj holds dense ids of the current
 * level, roots_cur maps them to level-1 root segments.  status |= 2 if the sweep cap was hit (the
 * reference loops forever in that case). */
int sgb_group_nearby(const int* adj, int A, const int* roots_cur, const float* dist, float th, int* uf, int S1,
                     int sweep_cap, int* status, void* stream);
int sgb_group_nearby_scenes(const int* adj, int A, const int* roots_cur, int S_cur, const float* dist, float th, int* uf, int S1,
                            int sweep_cap, int* status, const int* scene_seg_off, const int* scene_cl_off, int n_scenes,
                            int max_scene_segs, void* stream);

/* a14  one iteration of phase A of seggroup/model.py:439-470 `group_unlabeled_clusters`: row arg-min of
 * the dense distance matrix (fill 1000, first minimum), then union of every unlabeled cluster into it.
 * amin_ws: 2*S ints of scratch (arg-min row, then the in-order list of still unlabeled clusters). */
int sgb_group_unlabeled_step(const float* dist, const int* row_off, const int* nbr, const int* eid, int S,
                             const int* roots_cur, int* uf, int S1, int* amin_ws, void* stream);
/* scene batch: the dense distance matrix of model.py:312-316 is per scene, so a row's arg-min ranges over its own scene's
 * clusters (an isolated unlabeled cluster joins the FIRST cluster of its scene, as torch.min over an all-1000 row does). */
int sgb_group_unlabeled_step_scenes(const float* dist, const int* row_off, const int* nbr, const int* eid, int S,
                                    const int* roots_cur, int* uf, int S1, int* amin_ws,
                                    const int* scene_seg_off, const int* scene_cl_off, int n_scenes, int max_scene_segs,
                                    int max_scene_cl, void* stream);

/* a14  phase B of `group_unlabeled_clusters` (model.py:472-509): every cluster phase A left unlabeled (unl [n_unl], ascending
 * dense ids) joins the first labelled cluster of its candidate list cand [n_unl,S] (all cluster ids by increasing sampled-cloud
 * distance); further labelled candidates only receive the reference's stale-id point_num drift. */
int sgb_group_unlabeled_phase_b(const int* unl, int n_unl, const int* cand, int S, const int* roots_cur, int* uf, int S1,
                                void* stream);
/* a14  the candidate lists of phase B (model.py:472-487): cand [n_unl,width] = the clusters of the unlabeled cluster's own scene by
 * increasing min_p ||mean_i - p||^2 over the P sampled points of the candidate (cloud_idx [S,P] from sgb_cluster_cloud_indices),
 * ties -> lower id, rows padded with -1 (sgb_group_unlabeled_phase_b stops at the first -1). */
int sgb_phase_b_rank(const float* xyz, int stride, const int* cloud_idx, int P, const int* unl, int n_unl,
                     const int* scene_cl_off, int n_scenes, int S, int* cand, int width, void* stream);

/* a15  the classifier head and its loss, fused, one CTA per scene (forward and backward).
 * replaces seggroup/model.py:154-166 `Classifier.forward` (Linear 256->128 without bias, BatchNorm1d with BATCH statistics — the
 * reference never calls .eval() —, LeakyReLU 0.2, Dropout, Linear 128->40) on the per-instance features of model.py:902-921 and
 * seggroup/util.py:12-29 `cross_entropy_loss` (label smoothing 0.2, summed over the instances of the scene).
 * feat [G,256] instance features of all scenes, g_off [n_scenes+1] (device) instance range of every scene (>= 2 instances each:
 * torch's BatchNorm1d raises for one), gold [G] int32 target class, mask [G,128] 0/1 floats or NULL, drop_scale = 1/(1-p).
 * fwd out: hpre [G,128] (Linear1 output), stats [n_scenes,256] (batch mean, biased variance per scene), logits [G,40],
 * loss_raw [n_scenes,2] = (loss sum, instance count) as model.py:932 returns it.
 * bwd: grad_loss_sum [n_scenes]; scratch [G,128]; dfeat [G,256]; per-scene partial gradients dW1_part [n_scenes,128,256],
 * dgamma_part / dbeta_part [n_scenes,128], dW2_part [n_scenes,40,128], db2_part [n_scenes,40] (sum them in scene order). */
int sgb_classifier_head_fwd(const float* feat, int G, const int* g_off, int n_scenes, const int* gold,
                            const float* W1, const float* gamma, const float* beta, const float* W2, const float* b2,
                            const float* mask, float drop_scale, float* hpre, float* stats, float* logits, float* loss_raw, void* stream);
int sgb_classifier_head_bwd(const float* feat, int G, const int* g_off, int n_scenes, const int* gold,
                            const float* W1, const float* gamma, const float* beta, const float* W2,
                            const float* mask, float drop_scale, const float* hpre, const float* stats, const float* logits,
                            const float* grad_loss_sum, float* scratch, float* dfeat, float* dW1_part, float* dgamma_part,
                            float* dbeta_part, float* dW2_part, float* db2_part, void* stream);

/* a15  per-instance grouping of the final clusters, replaces seggroup/model.py:902-916 (`np.unique` of the clusters' weak instance
 * labels, the `torch.cat` of each group's clusters in ascending id order, the semantic label of the group's first cluster).
 * cl_ins / cl_sem [S] weak labels of the clusters (-1 = none: forms a group of its own, as np.unique does), scene_cl_off
 * [n_scenes+1] (device) cluster range of every scene (NULL with n_scenes = 1).  One launch, no host synchronisation.
 * out: order [S] cluster ids sorted by (scene, label, id); off [S+1] group offsets into `order` (G + 1 entries used);
 * gold [S] semantic label per group (G used); g_off [n_scenes+1] group range per scene; counts [2] = (G, min groups per scene). */
int sgb_classifier_groups(const int* cl_ins, const int* cl_sem, const int* scene_cl_off, int n_scenes, int S,
                          int* order, int* off, int* gold, int* g_off, int* counts, void* stream);

/* a16  replaces seggroup/model.py:525-605 `export_{segment,instance,semantic}_label` up to the text
 * formatting: per raw vertex r (p = unmap[r], int64 as stored in unmap.pth; NULL = identity):
 * seg = root point id of p's cluster, ins/sem = weak label + 1 or -1. */
int sgb_export_labels(const long long* unmap, int n_raw, const int* seg_of_point, const int* seg2cl, const int* cl_rootpt,
                      const int* cl_ins, const int* cl_sem, int* out_seg, int* out_ins, int* out_sem, void* stream);
/* scene batch: the segment label is the cluster's root point id INSIDE its scene (model.py:527-531) */
int sgb_export_labels_scenes(const long long* unmap, int n_raw, const int* seg_of_point, const int* seg2cl, const int* cl_rootpt,
                             const int* cl_ins, const int* cl_sem, int* out_seg, int* out_ins, int* out_sem,
                             const int* scene_pt_off, int n_scenes, void* stream);

/* a17  replaces seggroup/model.py:608-655 `evaluate`: 40-class intersection / union counts of the semantic and the
 * instance prediction and the four accuracies, over the raw vertices with real_label[:,0] != 0.
 * real_label [n,2] int64 (sem 1..40, ins), 16-byte aligned; sem_pred / ins_pred [n] int32 as sgb_export_labels writes them.
 * sem_valid_ids / ins_valid_ids: HOST arrays of class ids (1..63) = SEM_VALID_CLASS_IDS / INS_VALID_CLASS_IDS.
 * out [164] = IoU_sem [2][40], IoU_ins [2][40], acc [4] (fp32).  status: bit 64 is OR-ed in if a predicted id >= 65536. */
size_t sgb_evaluate_ws_bytes(void);
int sgb_evaluate(const long long* real_label, const int* sem_pred, const int* ins_pred, int n,
                 const int* sem_valid_ids, int n_sem_valid, const int* ins_valid_ids, int n_ins_valid,
                 float* out, int* status, void* ws, size_t ws_bytes, void* stream);
/* scene batch: raw_off [n_scenes+1] HOST array, raw-vertex range of every scene inside real_label / sem_pred / ins_pred;
 * out [n_scenes,164]; the scenes are evaluated one after the other on `stream` with the same workspace. */
int sgb_evaluate_scenes(const long long* real_label, const int* sem_pred, const int* ins_pred, const int* raw_off, int n_scenes,
                        const int* sem_valid_ids, int n_sem_valid, const int* ins_valid_ids, int n_ins_valid,
                        float* out, int* status, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a18  grid subsampling (boundaries B2 / B3)
 * replaces kpconv/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 (what
 * `grid_subsampling.compute`, wrapper.cpp:58-286, runs) and kpconv/tf_custom_ops/tf_subsampling/
 * grid_subsampling/grid_subsampling.cpp:5-150 (`GridSubsampling` / `BatchGridSubsampling` ops).
 * points [N,3]; feat [N,fdim] / cls [N,ldim] optional (NULL, dim 0); batches [B] device lengths (NULL = one cloud,
 * B = 1).  Voxel key, origin and NX/NY use the reference's fp32 expressions; barycentre = fp32 sum in input order
 * times (float)(1.0/count), feature mean = fp32 sum / (float)count (bit-identical to the reference).  Voxels are
 * listed per batch element in order of FIRST OCCURRENCE (the reference: hash-map iteration order); label ties ->
 * smallest label.  Outputs are sized for M = N; counts[0] <- M; out_first [N] = input index of each voxel's first
 * point (may be NULL); out_batches [B] (may be NULL).  status |= 4 bad batch lengths, |= 8 grid finer than 2^44 cells.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_grid_subsample_ws_bytes(int N, int B);
int sgb_grid_subsample(const float* xyz, const float* feat, const int* cls, int N, int fdim, int ldim,
                       const int* batches, int B, float dl, float* out_xyz, float* out_feat, int* out_cls,
                       int* out_first, int* out_batches, int* counts, int* status,
                       void* ws, size_t ws_bytes, void* stream);
/* per batch element min / max corner: minmax [B][6], boff_out [B+1] */
int sgb_batch_bounds(const float* xyz, int N, const int* batches, int B, float* minmax, int* boff_out, int* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a19  batched radius neighbours (boundary B3)
 * replaces kpconv/tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp:211-332 `batch_nanoflann_neighbors` (the
 * function the `BatchOrderedNeighbors` op calls, tf_batch_neighbors.cpp:93) and the brute-force variants :58-208.
 * Strict d2 < r*r with d2 = dx*dx + dy*dy + dz*dz in fp32 left to right; rows sorted by (d2, index), padded with Ns,
 * indices global (batch offset added).  The row width is the global maximum count -> two phases sharing `ws`:
 * _count builds the hashed support grid and writes the maximum count to max_count_out (device int), the caller
 * reads it back, allocates neighbors [Nq, W] and calls _fill.  status |= 16 grid too large, |= 32 > 1024 hits.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_radius_neighbors_ws_bytes(int Nq, int Ns, int B);
int sgb_radius_neighbors_count(const float* queries, int Nq, const float* supports, int Ns, const int* q_batches,
                               const int* s_batches, int B, float radius, int* max_count_out, int* status,
                               void* ws, size_t ws_bytes, void* stream);
int sgb_radius_neighbors_fill(const float* queries, int Nq, const float* supports, int Ns, int B, float radius, int W,
                              int* neighbors, int* status, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a20  rigid KPConv (boundary B4)
 * replaces kpconv/kernels/convolution_ops.py:161-249 `KPConv_ops` (argument order kept): query_points [n,3],
 * support_points [n0,3], neighbors [n,W] int32 (index >= n0 = shadow neighbour), features [n0,Cin], K_points [K,3],
 * K_values [K,Cin,Cout], KP_extent; influence 0 'linear' / 1 'constant' / 2 'gaussian' (sigma = 0.3 extent);
 * closest 0 'sum' / 1 'closest' aggregation.  out [n,Cout].  fp32 throughout (1e-4 relative vs an fp64 evaluation).
 * Limits: K <= 17, Cout multiple of 4 and <= 1024.
 * ------------------------------------------------------------------------------------------- */
int sgb_kpconv_fwd(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                   const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout, int K,
                   float KP_extent, int influence, int closest, float* out, void* stream);

/* gradients of sgb_kpconv_fwd w.r.t. features (gfeat [n0,Cin], accumulated: caller zero-fills) and K_values
 * (gK [K,Cin,Cout], overwritten; deterministic).  g [n,Cout] = upstream gradient. */
size_t sgb_kpconv_bwd_ws_bytes(int n, int Cin, int Cout, int K);
int sgb_kpconv_bwd(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                   const float* features, const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout,
                   int K, float KP_extent, int influence, int closest, float* gfeat, float* gK,
                   void* ws, size_t ws_bytes, void* stream);

/* the same gradients with both contractions on tcgen05 (TF32 x 3, fp32 accumulation in TMEM; kpconv_bwd_tc.cu):
 *   dK = WF^T g            persistent CTAs accumulate [K*Cin, Cout] in TMEM over all their 64-query tiles, partial sums reduced in a
 *                          fixed order (deterministic);
 *   dfeat += h_k * (g K_values[k]^T)   per 128-query tile and kernel point, scattered with coalesced red.global.add.
 * Gradient of kpconv/kernels/convolution_ops.py:240-247.  _supported: Cin in {32,64,128}, Cout a multiple of 32 up to 128, K <= 32,
 * W <= 64 (else use sgb_kpconv_bwd).  features, g and gfeat must be 16-byte aligned. */
int sgb_kpconv_bwd_tc_supported(int n, int W, int Cin, int Cout, int K, int n0);
size_t sgb_kpconv_bwd_tc_ws_bytes(int n, int Cin, int Cout, int K);
int sgb_kpconv_bwd_tc(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                      const float* features, const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout,
                      int K, float KP_extent, int influence, int closest, float* gfeat, float* gK,
                      void* ws, size_t ws_bytes, void* stream);

/* N3  deformable KPConv: replaces kpconv/kernels/convolution_ops.py:371-493 `KPConv_deform_ops`.
 * offsets [n,K,3]: per-query kernel-point displacements (the deformed kernel points are K_points + offsets[i]); modulations [n,K]
 * or NULL.  A neighbour farther than KP_extent from EVERY deformed kernel point is dropped (the reference compacts the neighbour
 * rows to the in-range ones, :426-445), 'constant' influence is the indicator d^2 < KP_extent^2, the per-kernel-point weighted
 * features are scaled by the modulations (:485-486).  With offsets == NULL these ARE sgb_kpconv_fwd / sgb_kpconv_bwd.
 * bwd additionally writes goffsets [n,K,3] and gmodulations [n,K] (overwritten). */
int sgb_kpconv_deform_fwd(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                          const float* K_points, const float* offsets, const float* modulations, const float* K_values,
                          int n, int n0, int W, int Cin, int Cout, int K, float KP_extent, int influence, int closest,
                          float* out, void* stream);
int sgb_kpconv_deform_bwd(const float* g, const float* query_points, const float* support_points, const int* neighbors,
                          const float* features, const float* K_points, const float* offsets, const float* modulations,
                          const float* K_values, int n, int n0, int W, int Cin, int Cout, int K, float KP_extent, int influence,
                          int closest, float* gfeat, float* gK, float* goffsets, float* gmodulations,
                          void* ws, size_t ws_bytes, void* stream);

/* N3  'fitting' regulariser of the deformable layers (kpconv/models/KPFCNN_model.py:242-266): arg [n,K] = column of the neighbour row
 * closest to every deformed kernel point (first minimum; shadow entries count as a point at 1000, convolution_ops.py:405). */
int sgb_deform_closest_neighbor(const float* query_points, const float* support_points, const int* neighbors, const float* K_points,
                                const float* offsets, int n, int n0, int W, int K, int* arg, void* stream);

/* a20 on the tensor cores: same contract as sgb_kpconv_fwd, with the [n, K*Cin] x [K*Cin, Cout] contraction of
 * convolution_ops.py:240-247 on tcgen05 (kind::tf32, TF32 x 3 split: fp32-level accuracy, 1e-4 relative holds).
 * Shapes: Cin multiple of 32, Cout multiple of 16 and <= 256, W <= 64, K <= 32 (sgb_kpconv_tc_supported -> 1);
 * anything else returns SGB_ERR_UNSUPPORTED and the caller uses sgb_kpconv_fwd.  `ws` holds the pre-split,
 * pre-swizzled image of K_values (sgb_kpconv_tc_ws_bytes). */
int sgb_kpconv_tc_supported(int W, int Cin, int Cout, int K, int n0);
size_t sgb_kpconv_tc_ws_bytes(int Cin, int Cout, int K);
int sgb_kpconv_fwd_tc(const float* query_points, const float* support_points, const int* neighbors, const float* features,
                      const float* K_points, const float* K_values, int n, int n0, int W, int Cin, int Cout, int K,
                      float KP_extent, int influence, int closest, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a21  index pooling beside KPConv (strided shortcut / upsampling blocks)
 * replaces kpconv/models/network_blocks.py:49-66 `ind_max_pool(x, inds)` and :69-81 `closest_pool(x, inds)`:
 * x [n1,d] f32, inds [n2,W] int32 with index >= n1 (or < 0) = shadow row.  ind_max_pool: out[i] = max over the W
 * listed rows, the shadow row being the column minimum of x (:58); closest_pool: out[i] = row inds[i,0], the
 * shadow row being zeros (:77-80).  Backward = the tf.reduce_max / tf.reduce_min / tf.gather gradients: ties share
 * the upstream gradient equally, the shadow share flows on to the rows attaining the column minimum.  gx [n1,d] is
 * overwritten.  `out` of the forward is an input of the backward.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_ind_max_pool_ws_bytes(int d);
int sgb_ind_max_pool_fwd(const float* x, int n1, int d, const int* inds, int n2, int W, float* out,
                         void* ws, size_t ws_bytes, void* stream);
int sgb_ind_max_pool_bwd(const float* g, const float* x, int n1, int d, const int* inds, int n2, int W,
                         const float* out, float* gx, void* ws, size_t ws_bytes, void* stream);
int sgb_closest_pool_fwd(const float* x, int n1, int d, const int* inds, int n2, int W, float* out, void* stream);
int sgb_closest_pool_bwd(const float* g, int n1, int d, const int* inds, int n2, int W, float* gx, void* stream);

/* HOST function: writes n lines '%d\n' to `path` — the text format of seggroup/model.py:536-546 that the stage-2
 * consumers read (kpconv/datasets/Scannet2.py:148-156).  `values` is a host pointer. */
int sgb_write_labels_host(const char* path, const int* values, int n);

/* ---------------------------------------------------------------------------------------------
 * N1  batch normalisation (+ residual) + LeakyReLU of the KPFCNN blocks
 * replaces kpconv/models/network_blocks.py:147-163 `batch_norm` (tf.layers.batch_normalization, epsilon 1e-6, batch statistics
 * over the points of the stacked batch in training) + 166-173 `leaky_relu`, and the residual join of the bottleneck blocks
 * (`leaky_relu(features + shortcut)`, 337, 581).
 *
 * y [n,d] = act(gamma * (x - mean) * invstd + beta + residual), act = LeakyReLU(slope) (slope = 1: identity).
 * training != 0: batch statistics; stat [3][d] <- mean, invstd, biased variance; running_mean / running_var (optional) are
 * updated as r <- (1 - momentum) r + momentum * batch value (unbiased variance).  training == 0: the running statistics.
 * gamma / beta / residual may be NULL.  ws: sgb_bn_act_ws_bytes(n, d).
 * Backward: dx, dres (gradient of the residual input, optional), dgamma, dbeta (optional) from dy, x, y, stat of the forward.
 * ------------------------------------------------------------------------------------------- */
size_t sgb_bn_act_ws_bytes(int n, int d);
int sgb_bn_act_fwd(const float* x, int n, int d, const float* gamma, const float* beta, const float* residual, float eps,
                   float slope, int training, float momentum, float* running_mean, float* running_var, float* y, float* stat,
                   void* ws, size_t ws_bytes, void* stream);
int sgb_bn_act_bwd(const float* dy, const float* x, const float* y, int n, int d, const float* gamma, const float* stat,
                   float slope, int training, float* dx, float* dres, float* dgamma, float* dbeta,
                   void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
