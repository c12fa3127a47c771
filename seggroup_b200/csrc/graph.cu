// a2/a3/a11/a12/a13/a14/a16: the segment graph — union-find state, level construction, adjacency
// update, edge distances, CSR GCN aggregation, order-exact grouping, label export.
// seggroup/model.py:169-258, 262-316, 439-509, 525-605.
//
// The reference keeps a POINT-level DisjointSet with explicit member lists; a cluster is always a
// concatenation of whole level-1 segments (model.py:191), so the device state is SEGMENT-level:
//   uf = int[6][S1]: parent, next, tail (member lists as linked lists of level-1 segments, the root is
//   the head of its list), pnum (point count), ins, sem (weak labels of the root).
// `union(a, b)` (model.py:181-192): veto if both labelled and different; root = b; b's list += a's list.
// Grouping is inherently sequential in edge order: it runs on ONE thread of one CTA
// with the state in L1/L2, after a parallel filter has discarded the edges that cannot matter.
#include "common.cuh"

namespace {
struct UF {
    int* parent; int* next; int* tail; int* pnum; int* ins; int* sem;
    __host__ __device__ UF(int* base, int S) : parent(base), next(base + S), tail(base + 2 * S), pnum(base + 3 * S),
                                               ins(base + 4 * S), sem(base + 5 * S) {}
    // rows of length n holding the entries [s0, s0 + n) of a table: indexed by the GLOBAL segment id (base - s0)
    __host__ __device__ UF(int* base, int n, int s0) : parent(base - s0), next(base + n - s0), tail(base + 2 * n - s0),
                                                       pnum(base + 3 * n - s0), ins(base + 4 * n - s0), sem(base + 5 * n - s0) {}
};

// Scene batches (block-diagonal graphs: several scenes concatenated, ids offset, no edge between scenes).  The order-dependent
// replays run one CTA per scene.  scene_seg_off [B+1]: level-1 segment range of every scene; scene_cl_off [B+1]: its cluster
// range at the level the kernel works on (clusters are numbered by ascending root segment, so a scene's clusters are
// contiguous).  NULL = one scene covering everything.
struct SceneRange { int s0, s1, c0, c1; };
__device__ __forceinline__ SceneRange scene_range(const int* __restrict__ scene_seg_off, const int* __restrict__ scene_cl_off, int b,
                                                  int S1, int S_cur) {
    SceneRange r;
    r.s0 = scene_seg_off ? scene_seg_off[b] : 0;  r.s1 = scene_seg_off ? scene_seg_off[b + 1] : S1;
    r.c0 = scene_cl_off ? scene_cl_off[b] : 0;    r.c1 = scene_cl_off ? scene_cl_off[b + 1] : S_cur;
    return r;
}
// first edge of a lexicographically sorted (u < v) edge list whose first endpoint is >= c
__device__ __forceinline__ int edge_lower_bound(const int* __restrict__ adj, int A, int c) {
    int lo = 0, hi = A;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (adj[2 * (size_t)mid] < c) lo = mid + 1; else hi = mid; }
    return lo;
}

__device__ __forceinline__ int uf_find(const UF& u, int s) {            // path halving (single writer)
    while (u.parent[s] != s) { u.parent[s] = u.parent[u.parent[s]]; s = u.parent[s]; }
    return s;
}
__device__ __forceinline__ int uf_find_ro(const int* __restrict__ parent, int s) {   // read-only, any thread
    int p = parent[s];
    while (p != s) { s = p; p = parent[s]; }
    return s;
}
__device__ __forceinline__ bool uf_union(const UF& u, int a, int b) {   // model.py:181-192
    if (a == b) return false;
    const int ia = u.ins[a], ib = u.ins[b];
    if (ia != -1 && ib != -1 && ia != ib) return false;
    u.parent[a] = b;
    u.pnum[b] += u.pnum[a];
    if (ia != ib) { u.ins[b] = -ia * ib; u.sem[b] = -u.sem[a] * u.sem[b]; }
    u.next[u.tail[b]] = a;
    u.tail[b] = u.tail[a];
    return true;
}

// ---------------------------------------------------------------------------------------------
// scene init (model.py:712-721)
// ---------------------------------------------------------------------------------------------
__global__ void scene_init_points(const int* __restrict__ seg_off, const int* __restrict__ seg_members, int N, int S,
                                  int* __restrict__ seg_of_point, int* __restrict__ seg_of_pos) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    const int s = sgb_upper_segment(seg_off, S, q);
    seg_of_pos[q] = s;
    seg_of_point[__ldg(seg_members + q)] = s;
}
__global__ void scene_init_uf(const int* __restrict__ seg_off, const int* __restrict__ seg_members, const int* __restrict__ weak,
                              int S, int* __restrict__ ufb) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    UF u(ufb, S);
    const int root_pt = __ldg(seg_members + __ldg(seg_off + s));
    u.parent[s] = s; u.next[s] = -1; u.tail[s] = s;
    u.pnum[s] = __ldg(seg_off + s + 1) - __ldg(seg_off + s);
    u.sem[s] = __ldg(weak + (size_t)root_pt * 2);
    u.ins[s] = __ldg(weak + (size_t)root_pt * 2 + 1);
}

// ---------------------------------------------------------------------------------------------
// level construction (model.py:209-214 get_cluster_list + the cluster_map dict loops 759-768)
// ---------------------------------------------------------------------------------------------
__global__ void level_flag_roots(const int* __restrict__ parent, int S, int* __restrict__ flag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) flag[s] = parent[s] == s ? 1 : 0;
}
__global__ void level_assign(const int* __restrict__ ufb, int S, const int* __restrict__ dense /*scan of flags, [S+1]*/,
                             const int* __restrict__ seg_off, const int* __restrict__ seg_members,
                             int* __restrict__ roots, int* __restrict__ seg2cl, int* __restrict__ cl_ins, int* __restrict__ cl_sem,
                             int* __restrict__ cl_rootpt, int* __restrict__ counts,
                             const int* __restrict__ scene_seg_off, int n_scenes, int* __restrict__ scene_cl_off) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) counts[0] = dense[S];
    if (scene_cl_off && s <= n_scenes) scene_cl_off[s] = dense[scene_seg_off[s]];   // clusters are numbered by ascending root segment
    if (s >= S) return;
    const int* parent = ufb;
    const int r = uf_find_ro(parent, s);
    const int c = dense[r];
    seg2cl[s] = c;
    if (r == s) {
        roots[c] = s;
        cl_ins[c] = ufb[4 * S + s];
        cl_sem[c] = ufb[5 * S + s];
        cl_rootpt[c] = __ldg(seg_members + __ldg(seg_off + s));
    }
}
// one thread per cluster walks its member list: segment order, per-segment start inside the cluster
__global__ void level_walk_lists(const int* __restrict__ ufb, int S, const int* __restrict__ counts, const int* __restrict__ roots,
                                 const int* __restrict__ seg_off, int* __restrict__ cl_nseg, int* __restrict__ cl_npt,
                                 int* __restrict__ seg_rank /*[S] index inside the list*/, int* __restrict__ seg_start /*[S] point offset inside the cluster*/) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = counts[0];
    if (c >= S) return;
    if (c >= nc) { cl_nseg[c] = 0; cl_npt[c] = 0; return; }
    const int* next = ufb + S;
    int s = roots[c], i = 0, pts = 0;
    while (s >= 0) {
        seg_rank[s] = i; seg_start[s] = pts;
        pts += __ldg(seg_off + s + 1) - __ldg(seg_off + s);
        ++i;
        s = next[s];
    }
    cl_nseg[c] = i; cl_npt[c] = pts;
}
__global__ void level_fill_seglist(int S, const int* __restrict__ seg2cl, const int* __restrict__ seg_rank,
                                   const int* __restrict__ cl_seg_off, int* __restrict__ cl_seg_list) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) cl_seg_list[cl_seg_off[seg2cl[s]] + seg_rank[s]] = s;
}
__global__ void level_fill_order(int N, const int* __restrict__ seg_of_pos, const int* __restrict__ seg_off,
                                 const int* __restrict__ seg_members, const int* __restrict__ seg2cl, const int* __restrict__ seg_start,
                                 const int* __restrict__ cl_pt_off, int* __restrict__ order) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    const int s = seg_of_pos[q];
    order[cl_pt_off[seg2cl[s]] + seg_start[s] + (q - __ldg(seg_off + s))] = __ldg(seg_members + q);
}

// children of every new cluster among the old clusters, ascending old index (model.py:760-768):
// one warp per new cluster scans the old->new map with ballots (ordered compaction, no atomics).
__global__ void children_count(const int* __restrict__ old2new, int n_old, int* __restrict__ cnt, int n_new_cap) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_old) atomicAdd(cnt + old2new[j], 1);
}
__global__ void children_fill(const int* __restrict__ old2new, int n_old, const int* __restrict__ child_off, int n_new,
                              int* __restrict__ child_list) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n_new) return;
    int w = child_off[c];
    const int end = child_off[c + 1];
    for (int j0 = 0; j0 < n_old && w < end; j0 += 32) {
        const int j = j0 + lane;
        const bool hit = j < n_old && old2new[j] == c;
        const unsigned m = __ballot_sync(SGB_FULL_MASK, hit);
        if (hit) child_list[w + __popc(m & ((1u << lane) - 1))] = j;
        w += __popc(m);
    }
}
__global__ void gather_old2new(const int* __restrict__ roots_old, int n_old, const int* __restrict__ seg2cl_new, int* __restrict__ old2new) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_old) old2new[j] = seg2cl_new[roots_old[j]];
}

// ---------------------------------------------------------------------------------------------
// update_adj (model.py:291-302): map, drop self edges, order each pair, unique-lexicographic.
// Dedupe through an S x S bitmap (S <= ~3k clusters -> <= 1.1 MB): set bits, then per-row ordered
// compaction.  Output is sorted by construction.
// ---------------------------------------------------------------------------------------------
__global__ void adj_set_bits(const int* __restrict__ edges, int E, const int* __restrict__ map, int S, int words_per_row,
                             unsigned* __restrict__ bitmap) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int a = map[__ldg(edges + 2 * (size_t)e)], b = map[__ldg(edges + 2 * (size_t)e + 1)];
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    atomicOr(bitmap + (size_t)a * words_per_row + (b >> 5), 1u << (b & 31));
}
__global__ void adj_row_count(const unsigned* __restrict__ bitmap, int S, int words_per_row, int* __restrict__ row_cnt) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= S) return;
    int c = 0;
    for (int w = lane; w < words_per_row; w += 32) c += __popc(bitmap[(size_t)r * words_per_row + w]);
    c = sgb_warp_sum(c);
    if (lane == 0) row_cnt[r] = c;
}
__global__ void adj_row_write(const unsigned* __restrict__ bitmap, int S, int words_per_row, const int* __restrict__ row_off,
                              int* __restrict__ adj_out, int* __restrict__ counts) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r == 0 && threadIdx.x == 0 && blockIdx.x == 0) counts[1] = row_off[S];
    if (r >= S) return;
    int w0 = row_off[r];
    for (int wb = 0; wb < words_per_row; wb += 32) {
        const int w = wb + lane;
        const unsigned bits = w < words_per_row ? bitmap[(size_t)r * words_per_row + w] : 0u;
        int c = __popc(bits);
        // exclusive prefix over lanes
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(SGB_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        int pos = w0 + inc - c;
        unsigned b = bits;
        while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1;
            adj_out[2 * (size_t)pos] = r;
            adj_out[2 * (size_t)pos + 1] = w * 32 + bit;
            ++pos;
        }
        w0 += __shfl_sync(SGB_FULL_MASK, inc, 31);
    }
}

// ---------------------------------------------------------------------------------------------
// symmetric CSR of a unique, lexicographically sorted (u < v) edge list, for gather-style reductions with a fixed
// summation order: row i lists (neighbour, edge id) by ascending neighbour.  No sort: a symmetric S x S bitmap
// (<= 1.1 MB at S = 3k, L2 resident) is filled from the edges and compacted in order, one warp per row; the edge id
// of (min, max) is first[min] + rank of max among the bits above the diagonal of row min, because adj is sorted.
// ---------------------------------------------------------------------------------------------
__global__ void csr_set_bits(const int* __restrict__ adj, int A, int wpr, unsigned* __restrict__ bitmap) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A) return;
    const int u = adj[2 * e], v = adj[2 * e + 1];
    atomicOr(bitmap + (size_t)u * wpr + (v >> 5), 1u << (v & 31));
    atomicOr(bitmap + (size_t)v * wpr + (u >> 5), 1u << (u & 31));
}
// bits of word w of row i strictly above column c
__device__ __forceinline__ unsigned bits_above(unsigned bits, int w, int c) {
    const int lo = c + 1 - w * 32;                     // first column of this word that counts
    if (lo <= 0) return bits;
    if (lo >= 32) return 0u;
    return bits & (0xffffffffu << lo);
}
__global__ void csr_row_degrees(const unsigned* __restrict__ bitmap, int S, int wpr, int* __restrict__ deg, int* __restrict__ outdeg) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= S) return;
    int d = 0, o = 0;
    for (int w = lane; w < wpr; w += 32) {
        const unsigned bits = bitmap[(size_t)r * wpr + w];
        d += __popc(bits);
        o += __popc(bits_above(bits, w, r));
    }
    d = sgb_warp_sum(d); o = sgb_warp_sum(o);
    if (lane == 0) { deg[r] = d; outdeg[r] = o; }
}
__global__ void csr_row_write(const unsigned* __restrict__ bitmap, int S, int wpr, const int* __restrict__ row_off,
                              const int* __restrict__ first, int* __restrict__ nbr, int* __restrict__ eid) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= S) return;
    const int base = row_off[r];
    const int indeg = (row_off[r + 1] - base) - (first[r + 1] - first[r]);
    const int first_r = first[r];
    int w0 = base;
    for (int wb = 0; wb < wpr; wb += 32) {
        const int w = wb + lane;
        const unsigned bits = w < wpr ? bitmap[(size_t)r * wpr + w] : 0u;
        const int c = __popc(bits);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(SGB_FULL_MASK, inc, o); if (lane >= o) inc += t; }
        int pos = w0 + inc - c;
        unsigned b = bits;
        while (b) {
            const int j = w * 32 + __ffs(b) - 1;
            b &= b - 1;
            int e;
            if (j > r) {
                e = first_r + (pos - base - indeg);
            } else {                                   // edge (j, r): rank of r among the bits above the diagonal of row j
                const unsigned* rowj = bitmap + (size_t)j * wpr;
                int rank = 0;
                for (int ww = j >> 5; ww <= (r >> 5); ++ww) {
                    unsigned bb = bits_above(rowj[ww], ww, j);
                    const int hi = r - ww * 32;        // keep columns < r
                    if (hi < 32) bb &= (hi <= 0 ? 0u : (0xffffffffu >> (32 - hi)));
                    rank += __popc(bb);
                }
                e = first[j] + rank;
            }
            nbr[pos] = j; eid[pos] = e;
            ++pos;
        }
        w0 += __shfl_sync(SGB_FULL_MASK, inc, 31);
    }
}

// ---------------------------------------------------------------------------------------------
// edge distances (model.py:269-274; F.pairwise_distance = ||a - b + 1e-6||_2): one warp per edge
// ---------------------------------------------------------------------------------------------
__global__ void edge_dist_fwd_kernel(const float* __restrict__ feat, int C, const int* __restrict__ adj, int A, float* __restrict__ dist) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= A) return;
    const float* a = feat + (size_t)adj[2 * e] * C;
    const float* b = feat + (size_t)adj[2 * e + 1] * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = (__ldg(a + c) - __ldg(b + c)) + 1e-6f; s = fmaf(d, d, s); }
    s = sgb_warp_sum(s);
    if (lane == 0) dist[e] = sqrtf(s);
}
// The three CSR row kernels below run one CTA per cluster.  A row's incident-edge metadata (edge id -> endpoints, distance,
// similarity, neighbour row sum) sits behind two or three DEPENDENT global loads; read inside the per-channel loop that
// chain is paid once per neighbour and a cluster with hundreds of neighbours (a floor) holds the whole launch for
// > 200 us.  The CTA therefore stages the row's metadata in shared memory cooperatively, CSR_CHUNK entries at a time, and
// the channel loop is left with one independent feature load per neighbour.  Summation order (ascending CSR position,
// diagonal in column order) is unchanged, so results are bit-identical to the straightforward loop.
constexpr int CSR_CHUNK = 256;

// grad_feat[i] += sum over incident edges of +-g[e] * (F[u]-F[v]+eps)/d[e]   (gather over the symmetric CSR)
__global__ void __launch_bounds__(128)
edge_dist_bwd_kernel(const float* __restrict__ feat, int C, const int* __restrict__ adj, const float* __restrict__ dist,
                     const float* __restrict__ gdist, const int* __restrict__ row_off, const int* __restrict__ eid,
                     int S, float* __restrict__ gfeat) {
    __shared__ int s_u[CSR_CHUNK], s_v[CSR_CHUNK];
    __shared__ float s_d[CSR_CHUNK], s_g[CSR_CHUNK];
    const int i = blockIdx.x;
    if (i >= S) return;
    const int a = row_off[i], b = row_off[i + 1];
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float acc = 0.f;
        for (int t0 = a; t0 < b; t0 += CSR_CHUNK) {
            const int n = min(CSR_CHUNK, b - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                const int e = eid[t0 + k];
                s_u[k] = adj[2 * e]; s_v[k] = adj[2 * e + 1]; s_d[k] = dist[e]; s_g[k] = gdist[e];
            }
            __syncthreads();
            if (c < C) {
#pragma unroll 4
                for (int k = 0; k < n; ++k) {
                    const int u = s_u[k], v = s_v[k];
                    const float d = s_d[k];
                    const float diff = (__ldg(feat + (size_t)u * C + c) - __ldg(feat + (size_t)v * C + c)) + 1e-6f;
                    const float val = d > 0.f ? s_g[k] * diff / d : 0.f;
                    acc += (u == i) ? val : -val;
                }
            }
        }
        if (c < C) gfeat[(size_t)i * C + c] += acc;
    }
}

// ---------------------------------------------------------------------------------------------
// GCN aggregation (model.py:305-309, 146-151): A = I + sym(sims); AX = (A / rowsum) X   as a CSR gather
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gcn_agg_fwd_kernel(const float* __restrict__ X, int C, const float* __restrict__ sims, const int* __restrict__ row_off,
                   const int* __restrict__ nbr, const int* __restrict__ eid, int S, float* __restrict__ AX,
                   float* __restrict__ rowsum) {
    __shared__ int s_j[CSR_CHUNK];
    __shared__ float s_w[CSR_CHUNK];
    __shared__ float s_rs;
    const int i = blockIdx.x;
    if (i >= S) return;
    const int a = row_off[i], b = row_off[i + 1];
    // row sum in CSR order (one thread adds, from staged similarities: a fixed, sequential order)
    float rs = 1.f;                                   // diagonal
    for (int t0 = a; t0 < b; t0 += CSR_CHUNK) {
        const int n = min(CSR_CHUNK, b - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) s_w[k] = sims[eid[t0 + k]];
        __syncthreads();
        if (threadIdx.x == 0) for (int k = 0; k < n; ++k) rs += s_w[k];
    }
    if (threadIdx.x == 0) { s_rs = rs; rowsum[i] = rs; }
    __syncthreads();
    rs = s_rs;
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float acc = 0.f;
        bool diag_done = false;
        for (int t0 = a; t0 < b; t0 += CSR_CHUNK) {
            const int n = min(CSR_CHUNK, b - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < n; k += blockDim.x) { s_j[k] = nbr[t0 + k]; s_w[k] = sims[eid[t0 + k]] / rs; }
            __syncthreads();
            if (c < C) {
#pragma unroll 4
                for (int k = 0; k < n; ++k) {         // ascending column order incl. the diagonal
                    const int j = s_j[k];
                    if (!diag_done && j > i) { acc = fmaf(1.f / rs, __ldg(X + (size_t)i * C + c), acc); diag_done = true; }
                    acc = fmaf(s_w[k], __ldg(X + (size_t)j * C + c), acc);
                }
            }
        }
        if (c < C) {
            if (!diag_done) acc = fmaf(1.f / rs, __ldg(X + (size_t)i * C + c), acc);
            AX[(size_t)i * C + c] = acc;
        }
    }
}
// dX[j] = sum_i A[i,j]/rs_i * dAX[i]  (structure symmetric: gather over row j, using the neighbour's rowsum)
__global__ void __launch_bounds__(128)
gcn_agg_bwd_x_kernel(const float* __restrict__ dAX, int C, const float* __restrict__ sims, const float* __restrict__ rowsum,
                     const int* __restrict__ row_off, const int* __restrict__ nbr, const int* __restrict__ eid, int S,
                     float* __restrict__ dX) {
    __shared__ int s_i[CSR_CHUNK];
    __shared__ float s_w[CSR_CHUNK];
    const int j = blockIdx.x;
    if (j >= S) return;
    const int a = row_off[j], b = row_off[j + 1];
    const float rsj = rowsum[j];
    for (int c0 = 0; c0 < C; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        float acc = c < C ? __ldg(dAX + (size_t)j * C + c) / rsj : 0.f;
        for (int t0 = a; t0 < b; t0 += CSR_CHUNK) {
            const int n = min(CSR_CHUNK, b - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                const int i = nbr[t0 + k];
                s_i[k] = i; s_w[k] = sims[eid[t0 + k]] / rowsum[i];
            }
            __syncthreads();
            if (c < C) {
#pragma unroll 4
                for (int k = 0; k < n; ++k) acc = fmaf(s_w[k], __ldg(dAX + (size_t)s_i[k] * C + c), acc);
            }
        }
        if (c < C) dX[(size_t)j * C + c] = acc;
    }
}
// dsims[e=(u,v)] = (dAX[u].X[v] - dAX[u].AX[u]) / rs_u + (dAX[v].X[u] - dAX[v].AX[v]) / rs_v : one warp per edge
__global__ void gcn_agg_bwd_s_kernel(const float* __restrict__ dAX, const float* __restrict__ X, const float* __restrict__ AX, int C,
                                     const float* __restrict__ rowsum, const int* __restrict__ adj, int A, float* __restrict__ dsims) {
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= A) return;
    const int u = adj[2 * e], v = adj[2 * e + 1];
    float s_uv = 0.f, s_uu = 0.f, s_vu = 0.f, s_vv = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float gu = __ldg(dAX + (size_t)u * C + c), gv = __ldg(dAX + (size_t)v * C + c);
        s_uv = fmaf(gu, __ldg(X + (size_t)v * C + c), s_uv);
        s_uu = fmaf(gu, __ldg(AX + (size_t)u * C + c), s_uu);
        s_vu = fmaf(gv, __ldg(X + (size_t)u * C + c), s_vu);
        s_vv = fmaf(gv, __ldg(AX + (size_t)v * C + c), s_vv);
    }
    s_uv = sgb_warp_sum(s_uv); s_uu = sgb_warp_sum(s_uu); s_vu = sgb_warp_sum(s_vu); s_vv = sgb_warp_sum(s_vv);
    if (lane == 0) dsims[e] = (s_uv - s_uu) / rowsum[u] + (s_vu - s_vv) / rowsum[v];
}

// ---------------------------------------------------------------------------------------------
// group_nearby_clusters (model.py:218-258) — sequential replay by ONE thread, everything else parallel:
// the union-find state is staged in shared memory (24 B per level-1 segment), edges stream through shared
// memory in chunks with their endpoints already resolved to level-1 roots and the `dist > th` test already
// applied by the whole CTA, and the small-cluster sweep only runs after a CTA-wide check found a cluster
// with < 5 points (a cluster never shrinks, so "no small endpoint now" == "the sweep would attempt nothing").
// status bit 1 (value 2): the sweep cap was hit (the reference would spin forever).
// ---------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_CHUNK = 2048;

// Ordered block-wide compaction: the threads hold one (keep, ru, rv) triple each for 256 consecutive edges; the kept pairs
// are appended to s_ru / s_rv at `base` in edge order.  Returns the new fill count (block-uniform).
__device__ __forceinline__ int gn_append(bool keep, int ru, int rv, int* s_ru, int* s_rv, int base, int* s_wcount) {
    const unsigned bal = __ballot_sync(SGB_FULL_MASK, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_wcount[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < GN_THREADS / 32; ++w) {
        const int c = s_wcount[w];
        before += w < warp ? c : 0;
        total += c;
    }
    if (keep) {
        const int o = base + before + __popc(bal & ((1u << lane) - 1u));
        s_ru[o] = ru; s_rv[o] = rv;
    }
    __syncthreads();
    return base + total;
}

// The unions are order dependent (label veto, root choice, member-list order), so they are replayed by ONE thread in edge
// order, on union-find state held in shared memory.  What the other 255 threads do is shrink that thread's work list: the
// edges of a chunk are filtered in parallel, in order, down to the ones that can act --
//   pass 1: dist <= th (NaN compares false -> merge attempt, as in Python);
//   pass 2: an endpoint cluster has < 5 points at the time the chunk is filtered.  Point counts only grow, so an edge
//           that fails the test then fails it for the rest of the sweep: the filter is exact, and thread 0 re-checks the
//           condition at replay time for the ones that passed.
__global__ void __launch_bounds__(GN_THREADS)
group_nearby_kernel(const int* __restrict__ adj_all, int A_all, const int* __restrict__ roots_cur, int S_cur,
                    const float* __restrict__ dist_all, float th, int* __restrict__ ufb, int S1,
                    int sweep_cap, int* __restrict__ status, int state_in_smem,
                    const int* __restrict__ scene_seg_off, const int* __restrict__ scene_cl_off) {
    extern __shared__ int gn_smem[];
    __shared__ int s_ru[GN_CHUNK], s_rv[GN_CHUNK];
    __shared__ int s_wcount[GN_THREADS / 32];
    __shared__ int s_flag;
    __shared__ int s_e0, s_e1;
    const SceneRange sr = scene_range(scene_seg_off, scene_cl_off, blockIdx.x, S1, S_cur);
    if (threadIdx.x == 0) {                           // one scene (no scene arrays): every edge
        s_e0 = scene_cl_off ? edge_lower_bound(adj_all, A_all, sr.c0) : 0;
        s_e1 = scene_cl_off ? edge_lower_bound(adj_all, A_all, sr.c1) : A_all;
    }
    const int n_loc = sr.s1 - sr.s0;
    if (state_in_smem) {
        for (int r = 0; r < 6; ++r)
            for (int i = threadIdx.x; i < n_loc; i += GN_THREADS) gn_smem[r * n_loc + i] = ufb[(size_t)r * S1 + sr.s0 + i];
    }
    __syncthreads();
    const int* adj = adj_all + 2 * (size_t)s_e0;
    const float* dist = dist_all + s_e0;
    const int A = s_e1 - s_e0;
    const UF u = state_in_smem ? UF(gn_smem, n_loc, sr.s0) : UF(ufb, S1);
    // pass 1
    for (int c0 = 0; c0 < A; c0 += GN_CHUNK) {
        const int n = min(GN_CHUNK, A - c0);
        int fill = 0;
        for (int i0 = 0; i0 < n; i0 += GN_THREADS) {
            const int e = c0 + i0 + threadIdx.x;
            bool keep = false;
            int ru = 0, rv = 0;
            if (i0 + threadIdx.x < n && !(dist[e] > th)) {
                ru = roots_cur[adj[2 * (size_t)e]]; rv = roots_cur[adj[2 * (size_t)e + 1]];
                keep = true;
            }
            fill = gn_append(keep, ru, rv, s_ru, s_rv, fill, s_wcount);
        }
        if (threadIdx.x == 0) {
            for (int i = 0; i < fill; ++i) uf_union(u, uf_find(u, s_ru[i]), uf_find(u, s_rv[i]));
        }
        __syncthreads();
    }
    // pass 2: sweeps over ALL edges while some endpoint cluster has < 5 points
    int sweeps = 0;
    while (true) {
        if (sweeps >= sweep_cap) {                    // would another sweep still find a small cluster?
            if (threadIdx.x == 0) s_flag = 0;
            __syncthreads();
            int any = 0;
            for (int e = threadIdx.x; e < A; e += GN_THREADS) {
                const int c1 = uf_find_ro(u.parent, roots_cur[adj[2 * e]]);
                const int c2 = uf_find_ro(u.parent, roots_cur[adj[2 * e + 1]]);
                any |= (u.pnum[c1] < 5 || u.pnum[c2] < 5);
            }
            if (any) s_flag = 1;
            __syncthreads();
            if (s_flag && threadIdx.x == 0) atomicOr(status, 2);
            break;
        }
        int attempts = 0;
        for (int c0 = 0; c0 < A; c0 += GN_CHUNK) {
            const int n = min(GN_CHUNK, A - c0);
            int fill = 0;
            for (int i0 = 0; i0 < n; i0 += GN_THREADS) {
                const int e = c0 + i0 + threadIdx.x;
                bool keep = false;
                int ru = 0, rv = 0;
                if (i0 + threadIdx.x < n) {
                    ru = roots_cur[adj[2 * (size_t)e]]; rv = roots_cur[adj[2 * (size_t)e + 1]];
                    keep = u.pnum[uf_find_ro(u.parent, ru)] < 5 || u.pnum[uf_find_ro(u.parent, rv)] < 5;
                }
                fill = gn_append(keep, ru, rv, s_ru, s_rv, fill, s_wcount);
            }
            attempts += fill;
            if (threadIdx.x == 0) {
                for (int i = 0; i < fill; ++i) {
                    const int c1 = uf_find(u, s_ru[i]), c2 = uf_find(u, s_rv[i]);
                    if (u.pnum[c1] < 5 || u.pnum[c2] < 5) uf_union(u, c1, c2);
                }
            }
            __syncthreads();
        }
        if (attempts == 0) break;                     // block-uniform: a sweep that found no small cluster ends the loop
        ++sweeps;
    }
    __syncthreads();
    if (state_in_smem) {
        for (int r = 0; r < 6; ++r)
            for (int i = threadIdx.x; i < n_loc; i += GN_THREADS) ufb[(size_t)r * S1 + sr.s0 + i] = gn_smem[r * n_loc + i];
    }
}

// group_unlabeled_clusters phase A, one iteration (model.py:441-470):
// row argmin of the dense distance matrix (fill 1000, first minimum) from the symmetric CSR, then the
// sequential unions of unlabeled clusters in ascending order.
__global__ void unlabeled_argmin_kernel(const float* __restrict__ dist, const int* __restrict__ row_off, const int* __restrict__ nbr,
                                        const int* __restrict__ eid, int S, int* __restrict__ amin,
                                        const int* __restrict__ scene_cl_off, int n_scenes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    // the dense matrix of the reference is per scene: its columns are this scene's clusters [c0, c1)
    int c0 = 0, c1 = S;
    if (scene_cl_off) {
        const int b = sgb_upper_segment(scene_cl_off, n_scenes, i);
        c0 = scene_cl_off[b]; c1 = scene_cl_off[b + 1];
    }
    const int a = row_off[i], b = row_off[i + 1];
    // dense row: dm[i][j] = dist(e) for neighbours, 1000 elsewhere (diagonal included)
    float best = INFINITY; int bj = -1;
    int expect = c0;                                  // smallest column not yet known to be a neighbour
    int first_fill = -1;
    for (int t = a; t < b; ++t) {
        const int j = nbr[t];
        if (first_fill < 0 && j > expect) first_fill = expect;
        if (j == expect) ++expect;
        else if (j > expect) expect = j + 1;
        const float d = dist[eid[t]];
        if (d < best) { best = d; bj = j; }
    }
    if (first_fill < 0 && expect < c1) first_fill = expect;
    // candidates: (best, bj) among edges, (1000, first_fill) among the filled entries; first minimum wins
    if (first_fill >= 0 && (bj < 0 || 1000.f < best || (1000.f == best && first_fill < bj))) bj = first_fill;
    amin[i] = bj < 0 ? c0 : bj;
}
// The ascending sequential unions of phase A.  One thread replays them; the union-find table, the root map and the argmin
// row live in shared memory for the replay (a dependent shared-memory access costs ~30 cycles, an L2 one ~10x that).
// A cluster that carries a label keeps it through every later union (model.py:188-190), so the clusters that are labelled
// when the kernel starts are filtered out in parallel, in order; thread 0 re-checks the rest at replay time.
constexpr int UU_THREADS = 256;
__global__ void __launch_bounds__(UU_THREADS)
unlabeled_union_kernel(const int* __restrict__ amin, int S_all, const int* __restrict__ roots_cur, int* __restrict__ ufb, int S1,
                       int state_in_smem, int* __restrict__ list_ws /*[S], used when the state stays in global memory*/,
                       const int* __restrict__ scene_seg_off, const int* __restrict__ scene_cl_off) {
    extern __shared__ int uu_smem[];
    __shared__ int s_wcount[UU_THREADS / 32];
    const SceneRange sr = scene_range(scene_seg_off, scene_cl_off, blockIdx.x, S1, S_all);
    const int n_loc = sr.s1 - sr.s0, S = sr.c1 - sr.c0;
    // cluster-indexed arrays of this scene, indexed by the GLOBAL dense id
    const int* s_roots = roots_cur;
    const int* s_amin = amin;
    int* s_list = list_ws + sr.c0;
    if (state_in_smem) {
        int* r = uu_smem + 6 * n_loc;
        int* a = r + S;
        s_list = a + S;
        for (int q = 0; q < 6; ++q)
            for (int i = threadIdx.x; i < n_loc; i += UU_THREADS) uu_smem[q * n_loc + i] = ufb[(size_t)q * S1 + sr.s0 + i];
        for (int i = threadIdx.x; i < S; i += UU_THREADS) { r[i] = roots_cur[sr.c0 + i]; a[i] = amin[sr.c0 + i]; }
        s_roots = r - sr.c0; s_amin = a - sr.c0;
    }
    __syncthreads();
    const UF u = state_in_smem ? UF(uu_smem, n_loc, sr.s0) : UF(ufb, S1);
    int fill = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = 0; i0 < S; i0 += UU_THREADS) {
        const int i = sr.c0 + i0 + threadIdx.x;
        const bool keep = i < sr.c1 && u.ins[uf_find_ro(u.parent, s_roots[i])] == -1;
        const unsigned bal = __ballot_sync(SGB_FULL_MASK, keep);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < UU_THREADS / 32; ++w) { const int c = s_wcount[w]; before += w < warp ? c : 0; total += c; }
        if (keep) s_list[fill + before + __popc(bal & ((1u << lane) - 1u))] = i;
        fill += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int t = 0; t < fill; ++t) {
            const int i = s_list[t];
            const int c1 = uf_find(u, s_roots[i]);
            if (u.ins[c1] != -1) continue;
            uf_union(u, c1, uf_find(u, s_roots[s_amin[i]]));
        }
    }
    __syncthreads();
    if (state_in_smem) {
        for (int q = 0; q < 6; ++q)
            for (int i = threadIdx.x; i < n_loc; i += UU_THREADS) ufb[(size_t)q * S1 + sr.s0 + i] = uu_smem[q * n_loc + i];
    }
}

// group_unlabeled_clusters phase B (model.py:472-509): every cluster that phase A left unlabeled walks the other clusters in
// order of increasing sampled-cloud distance (`cand` [n_unl][S], sorted by the caller on the device) and joins the first
// labelled one; the reference keeps calling union() with the now stale id for every further labelled candidate, which only
// drifts point_num -- reproduced.  Sequential by definition: one thread.
__global__ void unlabeled_phase_b_kernel(const int* __restrict__ unl, int n_unl, const int* __restrict__ cand, int S,
                                         const int* __restrict__ roots_cur, int* __restrict__ ufb, int S1) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    UF u(ufb, S1);
    for (int q = 0; q < n_unl; ++q) {
        const int i = unl[q];
        const int c1 = uf_find(u, roots_cur[i]);
        if (u.ins[c1] != -1) continue;
        bool merged = false;
        for (int t = 0; t < S; ++t) {
            const int j = cand[(size_t)q * S + t];
            if (j < 0) break;                          // end of this cluster's candidate list (scene batch: rows are padded)
            if (j == i) continue;
            const int c2 = uf_find(u, roots_cur[j]);
            if (u.ins[c2] == -1) continue;
            if (merged) { u.pnum[c2] += u.pnum[c1]; continue; }
            uf_union(u, c1, c2);
            merged = true;
        }
    }
}

// Candidate ranking of phase B (model.py:472-487): for the unlabeled cluster i, every cluster j of ITS scene ordered by
//     dmin(i, j) = min over the 1024 sampled points p of cluster j of ||mean_i - p||^2,   mean_i = mean of cluster i's 1024 samples
// (squared distance (dx^2 + dy^2) + dz^2 in fp32 without FMA contraction, ascending, ties -> lower cluster id).
// One CTA per unlabeled cluster: mean by a fixed-order tree, one warp per candidate cluster for the minimum, bitonic sort of
// (dmin, j) in shared memory.  cand row q: n_scene_clusters ids, then -1 up to `width`.
constexpr int PB_THREADS = 256;
constexpr int PB_MAX = 4096;                              // clusters of one scene at this stage (a few dozen in practice)
__global__ void __launch_bounds__(PB_THREADS)
phase_b_rank_kernel(const float* __restrict__ xyz, int stride, const int* __restrict__ cloud_idx, int P, const int* __restrict__ unl,
                    const int* __restrict__ scene_cl_off, int n_scenes, int S, int* __restrict__ cand, int width) {
    __shared__ float s_key[PB_MAX];
    __shared__ int s_id[PB_MAX];
    __shared__ float s_red[3][PB_THREADS];
    const int q = blockIdx.x, i = unl[q];
    int c0 = 0, c1 = S;
    if (scene_cl_off) { const int b = sgb_upper_segment(scene_cl_off, n_scenes, i); c0 = scene_cl_off[b]; c1 = scene_cl_off[b + 1]; }
    const int n = c1 - c0;
    // mean of the cluster's own samples: per-thread partial sums over a strided quarter, then a fixed binary tree
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int p = threadIdx.x; p < P; p += PB_THREADS) {
        const float* pt = xyz + (size_t)__ldg(cloud_idx + (size_t)i * P + p) * stride;
        ax += __ldg(pt); ay += __ldg(pt + 1); az += __ldg(pt + 2);
    }
    s_red[0][threadIdx.x] = ax; s_red[1][threadIdx.x] = ay; s_red[2][threadIdx.x] = az;
    __syncthreads();
    for (int o = PB_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
#pragma unroll
            for (int a = 0; a < 3; ++a) s_red[a][threadIdx.x] += s_red[a][threadIdx.x + o];
        }
        __syncthreads();
    }
    const float mx = s_red[0][0] / (float)P, my = s_red[1][0] / (float)P, mz = s_red[2][0] / (float)P;
    // minimum squared distance to every cluster of the scene: one warp per cluster
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int m = 1;
    while (m < n) m <<= 1;
    for (int t = threadIdx.x; t < m; t += PB_THREADS) { s_key[t] = INFINITY; s_id[t] = 0x7fffffff; }
    __syncthreads();
    for (int jj = warp; jj < n; jj += PB_THREADS / 32) {
        const int j = c0 + jj;
        float best = INFINITY;
        for (int p = lane; p < P; p += 32) {
            const float* pt = xyz + (size_t)__ldg(cloud_idx + (size_t)j * P + p) * stride;
            const float dx = __fsub_rn(mx, __ldg(pt)), dy = __fsub_rn(my, __ldg(pt + 1)), dz = __fsub_rn(mz, __ldg(pt + 2));
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            best = fminf(best, d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(SGB_FULL_MASK, best, o));
        if (lane == 0) { s_key[jj] = best; s_id[jj] = j; }
    }
    __syncthreads();
    // bitonic sort of (key, id) ascending
    for (int k = 2; k <= m; k <<= 1) {
        for (int jstep = k >> 1; jstep > 0; jstep >>= 1) {
            for (int t = threadIdx.x; t < m; t += PB_THREADS) {
                const int x = t ^ jstep;
                if (x > t) {
                    const bool up = (t & k) == 0;
                    const float ka = s_key[t], kb = s_key[x];
                    const int ia = s_id[t], ib = s_id[x];
                    const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
                    if (a_gt_b == up) { s_key[t] = kb; s_key[x] = ka; s_id[t] = ib; s_id[x] = ia; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < width; t += PB_THREADS) cand[(size_t)q * width + t] = t < n ? s_id[t] : -1;
}

// ---------------------------------------------------------------------------------------------
// label export (model.py:525-605): per raw vertex r, p = unmap[r]: segment = root point id of p's cluster,
// instance / semantic = weak label + 1 (or -1 when the cluster is unlabeled)
// ---------------------------------------------------------------------------------------------
// Four consecutive raw vertices per thread: the four unmap -> segment -> cluster -> label chains are independent, so a
// thread keeps four dependent-load chains in flight instead of one (the one-vertex version sat at 11-13 % of DRAM
// bandwidth on load latency), and the unmap reads / label writes are 16-byte accesses when the arrays are aligned.
constexpr int EXP_V = 4;
__global__ void __launch_bounds__(256)
export_labels_kernel(const long long* __restrict__ unmap, int n_raw, const int* __restrict__ seg_of_point,
                     const int* __restrict__ seg2cl, const int* __restrict__ cl_rootpt, const int* __restrict__ cl_ins,
                     const int* __restrict__ cl_sem, int* __restrict__ out_seg, int* __restrict__ out_ins, int* __restrict__ out_sem,
                     const int* __restrict__ scene_pt_off, int n_scenes) {
    const int r0 = (blockIdx.x * blockDim.x + threadIdx.x) * EXP_V;
    if (r0 >= n_raw) return;
    const bool full = r0 + EXP_V <= n_raw;
    int p[EXP_V];
    if (full && unmap && ((reinterpret_cast<uintptr_t>(unmap) & 15) == 0)) {
        const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(unmap + r0)), b = __ldg(reinterpret_cast<const longlong2*>(unmap + r0) + 1);
        p[0] = (int)a.x; p[1] = (int)a.y; p[2] = (int)b.x; p[3] = (int)b.y;
    } else {
#pragma unroll
        for (int v = 0; v < EXP_V; ++v) p[v] = r0 + v < n_raw ? (unmap ? (int)__ldg(unmap + r0 + v) : r0 + v) : 0;
    }
    int sg[EXP_V], c[EXP_V], seg[EXP_V], ins[EXP_V], sem[EXP_V];
#pragma unroll
    for (int v = 0; v < EXP_V; ++v) sg[v] = r0 + v < n_raw ? __ldg(seg_of_point + p[v]) : 0;
#pragma unroll
    for (int v = 0; v < EXP_V; ++v) c[v] = __ldg(seg2cl + sg[v]);
#pragma unroll
    for (int v = 0; v < EXP_V; ++v) {
        // scene batch: the segment label is the root point id INSIDE its scene
        const int base = scene_pt_off ? __ldg(scene_pt_off + sgb_upper_segment(scene_pt_off, n_scenes, p[v])) : 0;
        seg[v] = out_seg ? __ldg(cl_rootpt + c[v]) - base : 0;
        const int i_ = __ldg(cl_ins + c[v]), s_ = __ldg(cl_sem + c[v]);
        ins[v] = i_ != -1 ? i_ + 1 : -1;
        sem[v] = s_ != -1 ? s_ + 1 : -1;
    }
    auto put = [&](int* out, const int (&val)[EXP_V]) {
        if (!out) return;
        if (full && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
            *reinterpret_cast<int4*>(out + r0) = make_int4(val[0], val[1], val[2], val[3]);
        } else {
#pragma unroll
            for (int v = 0; v < EXP_V; ++v) if (r0 + v < n_raw) out[r0 + v] = val[v];
        }
    };
    put(out_seg, seg); put(out_ins, ins); put(out_sem, sem);
}
__global__ void count_unlabeled_kernel(const int* __restrict__ cl_ins, const int* __restrict__ counts, int* __restrict__ out) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int S = counts[0];
    int n = 0;
    for (int i = threadIdx.x; i < S; i += blockDim.x) n += cl_ins[i] == -1;
    n = sgb_warp_sum(n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) out[2] = s_n;
}

// ---------------------------------------------------------------------------------------------
// Small-graph path: one CTA per step.  With S <= SMALL_S clusters every phase of a level (flag -> scan -> assign -> walk ->
// scan -> fill ...) is a few thousand elements: as separate launches each phase costs a launch (~3 us of stream time, 27 of
// them per clustering level); as one 1024-thread CTA the phases are separated by __syncthreads and the scans run in place.
// The phase bodies are the kernels above, re-indexed by (threadIdx.x, blockDim.x).  Arrays written in one phase and read in
// a later one are passed WITHOUT const/__restrict__ so that no read goes through the non-coherent path.
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_S = 2048;
constexpr int SMALL_THREADS = 1024;

// out[0..n] = exclusive scan of in[0..n) (out[n] = total); in and out must not alias; all threads of the CTA call it
__device__ void block_scan_excl(const int* in, int n, int* out, int* s_warp /*[34] shared*/) {
    const int t = threadIdx.x, lane = t & 31, wp = t >> 5, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int b = min(n, t * per), e = min(n, b + per);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += in[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(SGB_FULL_MASK, inc, o); if (lane >= o) inc += v; }
    __syncthreads();                                   // s_warp may still be read by a previous call
    if (lane == 31) s_warp[wp] = inc;
    __syncthreads();
    if (wp == 0) {
        const int v = lane < (nt >> 5) ? s_warp[lane] : 0;
        int iv = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(SGB_FULL_MASK, iv, o); if (lane >= o) iv += u; }
        s_warp[lane] = iv - v;
        if (lane == 31) s_warp[32] = iv;
    }
    __syncthreads();
    int run = s_warp[wp] + inc - sum;
    for (int i = b; i < e; ++i) { out[i] = run; run += in[i]; }
    if (t == 0) out[n] = s_warp[32];
    __syncthreads();
}

__global__ void __launch_bounds__(SMALL_THREADS)
level_build_small(const int* __restrict__ ufb, int S, const int* __restrict__ seg_off, const int* __restrict__ seg_members,
                  int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* cl_ins, int* cl_sem, int* cl_rootpt,
                  int* counts, int* flag, int* dense, int* cl_nseg, int* cl_npt, int* seg_rank, int* seg_start,
                  const int* __restrict__ scene_seg_off, int n_scenes, int* scene_cl_off) {
    __shared__ int s_warp[34];
    __shared__ int s_unl;
    const int* parent = ufb;
    const int* next = ufb + S;
    if (threadIdx.x == 0) s_unl = 0;
    for (int s0 = threadIdx.x; s0 < S; s0 += blockDim.x) flag[s0] = parent[s0] == s0 ? 1 : 0;
    __syncthreads();
    block_scan_excl(flag, S, dense, s_warp);
    const int nc = dense[S];
    if (scene_cl_off)
        for (int b = threadIdx.x; b <= n_scenes; b += blockDim.x) scene_cl_off[b] = dense[scene_seg_off[b]];
    for (int s0 = threadIdx.x; s0 < S; s0 += blockDim.x) {                // level_assign
        const int r = uf_find_ro(parent, s0);
        const int c = dense[r];
        seg2cl[s0] = c;
        if (r == s0) {
            roots[c] = s0;
            cl_ins[c] = ufb[4 * S + s0];
            cl_sem[c] = ufb[5 * S + s0];
            cl_rootpt[c] = __ldg(seg_members + __ldg(seg_off + s0));
        }
    }
    __syncthreads();
    int unl = 0;
    for (int c = threadIdx.x; c < S; c += blockDim.x) {                   // level_walk_lists (+ count of unlabeled clusters)
        if (c >= nc) { cl_nseg[c] = 0; cl_npt[c] = 0; continue; }
        int sgm = roots[c], i = 0, pts = 0;
        while (sgm >= 0) {
            seg_rank[sgm] = i; seg_start[sgm] = pts;
            pts += __ldg(seg_off + sgm + 1) - __ldg(seg_off + sgm);
            ++i;
            sgm = next[sgm];
        }
        cl_nseg[c] = i; cl_npt[c] = pts;
        unl += cl_ins[c] == -1;
    }
    unl = sgb_warp_sum(unl);
    if ((threadIdx.x & 31) == 0 && unl) atomicAdd(&s_unl, unl);
    __syncthreads();
    block_scan_excl(cl_nseg, S, cl_seg_off, s_warp);
    block_scan_excl(cl_npt, S, cl_pt_off, s_warp);
    for (int s0 = threadIdx.x; s0 < S; s0 += blockDim.x) cl_seg_list[cl_seg_off[seg2cl[s0]] + seg_rank[s0]] = s0;   // level_fill_seglist
    if (threadIdx.x == 0) { counts[0] = nc; counts[2] = s_unl; }
}

__global__ void __launch_bounds__(SMALL_THREADS)
children_small(const int* __restrict__ roots_old, int n_old, const int* __restrict__ seg2cl_new, int n_new,
               int* old2new, int* child_off, int* child_list, int* cnt) {
    __shared__ int s_warp[34];
    for (int c = threadIdx.x; c <= n_new; c += blockDim.x) cnt[c] = 0;
    for (int j = threadIdx.x; j < n_old; j += blockDim.x) old2new[j] = seg2cl_new[roots_old[j]];
    __syncthreads();
    for (int j = threadIdx.x; j < n_old; j += blockDim.x) atomicAdd(cnt + old2new[j], 1);
    __syncthreads();
    block_scan_excl(cnt, n_new, child_off, s_warp);
    // ordered fill in O(n_old): ONE warp walks the old clusters in ascending order, 32 at a time; lanes with the same new
    // cluster (match_any) take consecutive slots behind that cluster's cursor, so every child list comes out ascending
    // (the multi-CTA children_fill scans all of old2new once per new cluster: fine on 148 SMs, quadratic inside one CTA)
    for (int c = threadIdx.x; c < n_new; c += blockDim.x) cnt[c] = child_off[c];
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        volatile int* cursor = cnt;
        for (int j0 = 0; j0 < n_old; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < n_old;
            const int dest = valid ? old2new[j] : -1 - lane;
            const unsigned m = __match_any_sync(SGB_FULL_MASK, dest);
            const int rank = __popc(m & ((1u << lane) - 1u));
            const int base = valid ? cursor[dest] : 0;
            __syncwarp();
            if (valid) {
                child_list[base + rank] = j;
                if (rank == 0) cursor[dest] = base + __popc(m);
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS)
adj_rows_small(const unsigned* bitmap, int S, int wpr, int* row_cnt, int* row_off, int* adj_out, int* counts) {
    __shared__ int s_warp[34];
    const int lane = threadIdx.x & 31;
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {         // adj_row_count
        int c = 0;
        for (int w = lane; w < wpr; w += 32) c += __popc(bitmap[(size_t)r * wpr + w]);
        c = sgb_warp_sum(c);
        if (lane == 0) row_cnt[r] = c;
    }
    __syncthreads();
    block_scan_excl(row_cnt, S, row_off, s_warp);
    if (threadIdx.x == 0) counts[1] = row_off[S];
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {         // adj_row_write
        int w0 = row_off[r];
        for (int wb = 0; wb < wpr; wb += 32) {
            const int w = wb + lane;
            const unsigned bits = w < wpr ? bitmap[(size_t)r * wpr + w] : 0u;
            const int c = __popc(bits);
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(SGB_FULL_MASK, inc, o); if (lane >= o) inc += t; }
            int pos = w0 + inc - c;
            unsigned b = bits;
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                adj_out[2 * (size_t)pos] = r;
                adj_out[2 * (size_t)pos + 1] = w * 32 + bit;
                ++pos;
            }
            w0 += __shfl_sync(SGB_FULL_MASK, inc, 31);
        }
    }
}

__global__ void __launch_bounds__(SMALL_THREADS)
csr_degrees_small(const unsigned* bitmap, int S, int wpr, int has_edges, int* deg, int* outdeg, int* row_off, int* first) {
    __shared__ int s_warp[34];
    const int lane = threadIdx.x & 31;
    for (int r = threadIdx.x >> 5; r < S; r += blockDim.x >> 5) {         // csr_row_degrees
        int d = 0, o = 0;
        for (int w = lane; w < wpr; w += 32) {
            const unsigned bits = bitmap[(size_t)r * wpr + w];
            d += __popc(bits);
            o += __popc(bits_above(bits, w, r));
        }
        d = sgb_warp_sum(d); o = sgb_warp_sum(o);
        if (lane == 0) { deg[r] = d; outdeg[r] = o; }
    }
    __syncthreads();
    block_scan_excl(deg, S, row_off, s_warp);
    if (!has_edges) return;
    block_scan_excl(outdeg, S, first, s_warp);
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" int sgb_scene_init(const int* seg_off, const int* seg_members, const int* weak_label, int N, int S,
                              int* seg_of_point, int* seg_of_pos, int* uf, void* stream) {
    if (N <= 0 || S <= 0 || !seg_off || !seg_members || !weak_label || !seg_of_point || !seg_of_pos || !uf) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    { scene_init_points<<<sgb_div_up(N, 256), 256, 0, st>>>(seg_off, seg_members, N, S, seg_of_point, seg_of_pos); SGB_COUNT_LAUNCH(); }
    { scene_init_uf<<<sgb_div_up(S, 256), 256, 0, st>>>(seg_off, seg_members, weak_label, S, uf); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" size_t sgb_level_ws_bytes(int S1) { return (size_t)(6 * (S1 + 1)) * sizeof(int) + sgb_scan_ws_bytes(S1 + 1); }

// counts[0] <- number of clusters.  All outputs are sized for the level-1 segment count S1 (cl_*_off: S1+1).
extern "C" int sgb_level_build(const int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                               int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                               int* cl_ins, int* cl_sem, int* cl_rootpt, int* counts, void* ws, size_t ws_bytes, void* stream) {
    return sgb_level_build_scenes(uf, S1, N, seg_off, seg_members, seg_of_pos, roots, seg2cl, cl_seg_off, cl_seg_list, cl_pt_off, order,
                                  cl_ins, cl_sem, cl_rootpt, counts, nullptr, 1, nullptr, ws, ws_bytes, stream);
}
// scene batch: scene_seg_off [n_scenes+1] (device) -> scene_cl_off [n_scenes+1] (device): cluster range of every scene
extern "C" int sgb_level_build_scenes(const int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                                      int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                                      int* cl_ins, int* cl_sem, int* cl_rootpt, int* counts,
                                      const int* scene_seg_off, int n_scenes, int* scene_cl_off,
                                      void* ws, size_t ws_bytes, void* stream) {
    if (scene_cl_off && (!scene_seg_off || n_scenes < 1)) return SGB_ERR_INVALID;
    if (S1 <= 0 || N <= 0 || !uf || !seg_off || !seg_members || !seg_of_pos || !roots || !seg2cl || !cl_seg_off || !cl_seg_list ||
        !cl_pt_off || !order || !cl_ins || !cl_sem || !cl_rootpt || !counts || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_level_ws_bytes(S1)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* flag = (int*)ws;                 // [S1+1]
    int* dense = flag + (S1 + 1);         // [S1+1]
    int* cl_nseg = dense + (S1 + 1);      // [S1+1]
    int* cl_npt = cl_nseg + (S1 + 1);     // [S1+1]
    int* seg_rank = cl_npt + (S1 + 1);    // [S1+1]
    int* seg_start = seg_rank + (S1 + 1); // [S1+1]
    void* scan_ws = seg_start + (S1 + 1);
    const size_t scan_bytes = sgb_scan_ws_bytes(S1 + 1);
    const int g = sgb_div_up(S1, 256);
    int rc;
    if (S1 <= SMALL_S) {                  // one CTA for every cluster-sized phase (counts[0] = clusters, counts[2] = unlabeled ones)
        { level_build_small<<<1, SMALL_THREADS, 0, st>>>(uf, S1, seg_off, seg_members, roots, seg2cl, cl_seg_off, cl_seg_list, cl_pt_off,
                                                         cl_ins, cl_sem, cl_rootpt, counts, flag, dense, cl_nseg, cl_npt, seg_rank, seg_start,
                                                         scene_seg_off, n_scenes, scene_cl_off); SGB_COUNT_LAUNCH(); }
        { level_fill_order<<<sgb_div_up(N, 256), 256, 0, st>>>(N, seg_of_pos, seg_off, seg_members, seg2cl, seg_start, cl_pt_off, order); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
    { level_flag_roots<<<g, 256, 0, st>>>(uf, S1, flag); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(flag, dense, S1, scan_ws, scan_bytes, st))) return rc;
    { level_assign<<<g, 256, 0, st>>>(uf, S1, dense, seg_off, seg_members, roots, seg2cl, cl_ins, cl_sem, cl_rootpt, counts,
                                      scene_seg_off, n_scenes, scene_cl_off); SGB_COUNT_LAUNCH(); }
    { level_walk_lists<<<g, 256, 0, st>>>(uf, S1, counts, roots, seg_off, cl_nseg, cl_npt, seg_rank, seg_start); SGB_COUNT_LAUNCH(); }
    if ((rc = sgb_exclusive_scan_i32(cl_nseg, cl_seg_off, S1, scan_ws, scan_bytes, st))) return rc;
    if ((rc = sgb_exclusive_scan_i32(cl_npt, cl_pt_off, S1, scan_ws, scan_bytes, st))) return rc;
    { level_fill_seglist<<<g, 256, 0, st>>>(S1, seg2cl, seg_rank, cl_seg_off, cl_seg_list); SGB_COUNT_LAUNCH(); }
    { level_fill_order<<<sgb_div_up(N, 256), 256, 0, st>>>(N, seg_of_pos, seg_off, seg_members, seg2cl, seg_start, cl_pt_off, order); SGB_COUNT_LAUNCH(); }
    { count_unlabeled_kernel<<<1, 256, 0, st>>>(cl_ins, counts, counts); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// children CSR: child_off [n_new+1], child_list [n_old] (old dense cluster ids, ascending inside each new cluster),
// old2new [n_old].  ws: (n_new + 1) ints + scan workspace.
extern "C" size_t sgb_children_ws_bytes(int n_new) { return (size_t)(n_new + 1) * sizeof(int) + sgb_scan_ws_bytes(n_new + 1); }
extern "C" int sgb_level_children(const int* roots_old, int n_old, const int* seg2cl_new, int n_new, int* old2new,
                                  int* child_off, int* child_list, void* ws, size_t ws_bytes, void* stream) {
    if (n_old <= 0 || n_new <= 0 || !roots_old || !seg2cl_new || !old2new || !child_off || !child_list || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_children_ws_bytes(n_new)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int* cnt = (int*)ws;
    void* scan_ws = cnt + (n_new + 1);
    if (n_old <= SMALL_S && n_new <= SMALL_S) {
        { children_small<<<1, SMALL_THREADS, 0, st>>>(roots_old, n_old, seg2cl_new, n_new, old2new, child_off, child_list, cnt); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
    SGB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n_new + 1) * sizeof(int), st));
    { gather_old2new<<<sgb_div_up(n_old, 256), 256, 0, st>>>(roots_old, n_old, seg2cl_new, old2new); SGB_COUNT_LAUNCH(); }
    { children_count<<<sgb_div_up(n_old, 256), 256, 0, st>>>(old2new, n_old, cnt, n_new); SGB_COUNT_LAUNCH(); }
    int rc;
    if ((rc = sgb_exclusive_scan_i32(cnt, child_off, n_new, scan_ws, sgb_scan_ws_bytes(n_new + 1), st))) return rc;
    { children_fill<<<sgb_div_up(n_new, 8), 256, 0, st>>>(old2new, n_old, child_off, n_new, child_list); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" size_t sgb_update_adj_ws_bytes(int S_new) {
    const size_t wpr = (size_t)(S_new + 31) / 32;
    return (size_t)S_new * wpr * 4 + (size_t)(2 * (S_new + 1)) * sizeof(int) + sgb_scan_ws_bytes(S_new + 1);
}
// edges [E,2] (ids of the old level), map [n_old] -> new dense ids (< S_new).  adj_out capacity must be >= min(E, S_new*(S_new-1)/2)
// rows; counts[1] <- number of rows written.
extern "C" int sgb_update_adj(const int* edges, int E, const int* map, int S_new, int* adj_out, int* counts,
                              void* ws, size_t ws_bytes, void* stream) {
    if (E < 0 || S_new <= 0 || !map || !adj_out || !counts || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_update_adj_ws_bytes(S_new)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = (S_new + 31) / 32;
    unsigned* bitmap = (unsigned*)ws;
    int* row_cnt = (int*)(bitmap + (size_t)S_new * wpr);
    int* row_off = row_cnt + (S_new + 1);
    void* scan_ws = row_off + (S_new + 1);
    SGB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)S_new * wpr * 4, st));
    if (E > 0) {
        if (!edges) return SGB_ERR_INVALID;
        { adj_set_bits<<<sgb_div_up(E, 256), 256, 0, st>>>(edges, E, map, S_new, wpr, bitmap); SGB_COUNT_LAUNCH(); }
    }
    if (S_new <= SMALL_S) {
        { adj_rows_small<<<1, SMALL_THREADS, 0, st>>>(bitmap, S_new, wpr, row_cnt, row_off, adj_out, counts); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
    { adj_row_count<<<sgb_div_up(S_new, 8), 256, 0, st>>>(bitmap, S_new, wpr, row_cnt); SGB_COUNT_LAUNCH(); }
    int rc;
    if ((rc = sgb_exclusive_scan_i32(row_cnt, row_off, S_new, scan_ws, sgb_scan_ws_bytes(S_new + 1), st))) return rc;
    { adj_row_write<<<sgb_div_up(S_new, 8), 256, 0, st>>>(bitmap, S_new, wpr, row_off, adj_out, counts); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" size_t sgb_sym_csr_ws_bytes(int S) {
    const size_t wpr = (size_t)(S + 31) / 32;
    return (size_t)S * wpr * 4 + (size_t)(3 * (S + 1)) * sizeof(int) + sgb_scan_ws_bytes(S + 1);
}
extern "C" int sgb_sym_csr(const int* adj, int A, int S, int* row_off, int* nbr, int* eid, void* ws, size_t ws_bytes, void* stream) {
    if (A < 0 || S <= 0 || !row_off || !ws) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_sym_csr_ws_bytes(S)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = (S + 31) / 32;
    unsigned* bitmap = (unsigned*)ws;
    int* deg = (int*)(bitmap + (size_t)S * wpr);
    int* outdeg = deg + (S + 1);
    int* first = outdeg + (S + 1);
    void* scan_ws = first + (S + 1);
    const size_t scan_bytes = sgb_scan_ws_bytes(S + 1);
    SGB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)S * wpr * 4, st));
    if (A > 0) {
        if (!adj || !nbr || !eid) return SGB_ERR_INVALID;
        { csr_set_bits<<<sgb_div_up(A, 256), 256, 0, st>>>(adj, A, wpr, bitmap); SGB_COUNT_LAUNCH(); }
    }
    if (S <= SMALL_S) {
        // degrees + both scans in one CTA; the row write ranks every lower-triangle edge inside its partner row
        // (up to wpr words per edge), which wants all the SMs: one warp per row
        { csr_degrees_small<<<1, SMALL_THREADS, 0, st>>>(bitmap, S, wpr, A > 0 ? 1 : 0, deg, outdeg, row_off, first); SGB_COUNT_LAUNCH(); }
        if (A > 0) { csr_row_write<<<sgb_div_up(S, 8), 256, 0, st>>>(bitmap, S, wpr, row_off, first, nbr, eid); SGB_COUNT_LAUNCH(); }
        SGB_CHECK_LAUNCH();
        return SGB_OK;
    }
    { csr_row_degrees<<<sgb_div_up(S, 8), 256, 0, st>>>(bitmap, S, wpr, deg, outdeg); SGB_COUNT_LAUNCH(); }
    int rc;
    if ((rc = sgb_exclusive_scan_i32(deg, row_off, S, scan_ws, scan_bytes, st))) return rc;
    if (A > 0) {
        if ((rc = sgb_exclusive_scan_i32(outdeg, first, S, scan_ws, scan_bytes, st))) return rc;
        { csr_row_write<<<sgb_div_up(S, 8), 256, 0, st>>>(bitmap, S, wpr, row_off, first, nbr, eid); SGB_COUNT_LAUNCH(); }
    }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_edge_dist_fwd(const float* feat, int C, const int* adj, int A, float* dist, void* stream) {
    if (A < 0 || C <= 0) return SGB_ERR_INVALID;
    if (A == 0) return SGB_OK;
    if (!feat || !adj || !dist) return SGB_ERR_INVALID;
    { edge_dist_fwd_kernel<<<sgb_div_up(A, 8), 256, 0, (cudaStream_t)stream>>>(feat, C, adj, A, dist); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
// gfeat [S,C] += d(dist)/d(feat) * gdist  (gfeat must be initialised by the caller)
extern "C" int sgb_edge_dist_bwd(const float* feat, int S, int C, const int* adj, int A, const float* dist, const float* gdist,
                                 const int* row_off, const int* eid, float* gfeat, void* stream) {
    if (A < 0 || C <= 0 || S <= 0) return SGB_ERR_INVALID;
    if (A == 0) return SGB_OK;
    if (!feat || !adj || !dist || !gdist || !row_off || !eid || !gfeat) return SGB_ERR_INVALID;
    { edge_dist_bwd_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(feat, C, adj, dist, gdist, row_off, eid, S, gfeat); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_gcn_agg_fwd(const float* X, int S, int C, const float* sims, const int* row_off, const int* nbr, const int* eid,
                               float* AX, float* rowsum, void* stream) {
    if (S <= 0 || C <= 0 || !X || !row_off || !AX || !rowsum) return SGB_ERR_INVALID;
    { gcn_agg_fwd_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(X, C, sims, row_off, nbr, eid, S, AX, rowsum); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
extern "C" int sgb_gcn_agg_bwd(const float* dAX, const float* X, const float* AX, int S, int C, const float* sims, const float* rowsum,
                               const int* adj, int A, const int* row_off, const int* nbr, const int* eid,
                               float* dX, float* dsims, void* stream) {
    if (S <= 0 || C <= 0 || A < 0 || !dAX || !X || !AX || !rowsum || !row_off || !dX) return SGB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    { gcn_agg_bwd_x_kernel<<<S, 128, 0, st>>>(dAX, C, sims, rowsum, row_off, nbr, eid, S, dX); SGB_COUNT_LAUNCH(); }
    if (A > 0) {
        if (!adj || !dsims || !sims) return SGB_ERR_INVALID;
        { gcn_agg_bwd_s_kernel<<<sgb_div_up(A, 8), 256, 0, st>>>(dAX, X, AX, C, rowsum, adj, A, dsims); SGB_COUNT_LAUNCH(); }
    }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_group_nearby(const int* adj, int A, const int* roots_cur, const float* dist, float th, int* uf, int S1,
                                int sweep_cap, int* status, void* stream) {
    return sgb_group_nearby_scenes(adj, A, roots_cur, 0, dist, th, uf, S1, sweep_cap, status, nullptr, nullptr, 1, S1, stream);
}
// scene batch: one CTA per scene replays that scene's edges.  scene_seg_off / scene_cl_off [n_scenes+1] (device; cluster offsets of
// the level `adj` / `roots_cur` belong to), max_scene_segs = largest level-1 segment count of a scene (sizes the shared memory).
extern "C" int sgb_group_nearby_scenes(const int* adj, int A, const int* roots_cur, int S_cur, const float* dist, float th, int* uf, int S1,
                                       int sweep_cap, int* status, const int* scene_seg_off, const int* scene_cl_off, int n_scenes,
                                       int max_scene_segs, void* stream) {
    if (A < 0 || S1 <= 0 || !uf || !status || n_scenes < 1) return SGB_ERR_INVALID;
    if (A == 0) return SGB_OK;
    if (!adj || !roots_cur || !dist) return SGB_ERR_INVALID;
    if (n_scenes > 1 && (!scene_seg_off || !scene_cl_off)) return SGB_ERR_INVALID;
    if (!scene_seg_off) { n_scenes = 1; max_scene_segs = S1; }
    const size_t state_bytes = (size_t)6 * max_scene_segs * sizeof(int);
    const int in_smem = state_bytes <= 160 * 1024;
    if (in_smem && state_bytes > 16 * 1024)
        SGB_OPT_IN_SMEM(group_nearby_kernel);
    { group_nearby_kernel<<<n_scenes, GN_THREADS, in_smem ? state_bytes : 0, (cudaStream_t)stream>>>(
          adj, A, roots_cur, S_cur, dist, th, uf, S1, sweep_cap, status, in_smem, scene_seg_off, scene_cl_off); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// one iteration of phase A of group_unlabeled_clusters; amin_ws [2*S] ints
extern "C" int sgb_group_unlabeled_step(const float* dist, const int* row_off, const int* nbr, const int* eid, int S,
                                        const int* roots_cur, int* uf, int S1, int* amin_ws, void* stream) {
    return sgb_group_unlabeled_step_scenes(dist, row_off, nbr, eid, S, roots_cur, uf, S1, amin_ws, nullptr, nullptr, 1, S1, S, stream);
}
// scene batch: the dense distance matrix of the reference is per scene (arg-min columns = the scene's own clusters; an isolated
// unlabeled cluster joins the FIRST cluster of its scene); max_scene_cl = largest cluster count of a scene at this level.
extern "C" int sgb_group_unlabeled_step_scenes(const float* dist, const int* row_off, const int* nbr, const int* eid, int S,
                                               const int* roots_cur, int* uf, int S1, int* amin_ws,
                                               const int* scene_seg_off, const int* scene_cl_off, int n_scenes, int max_scene_segs,
                                               int max_scene_cl, void* stream) {
    if (S <= 0 || S1 <= 0 || !row_off || !roots_cur || !uf || !amin_ws || n_scenes < 1) return SGB_ERR_INVALID;
    if (n_scenes > 1 && (!scene_seg_off || !scene_cl_off)) return SGB_ERR_INVALID;
    if (!scene_seg_off) { n_scenes = 1; max_scene_segs = S1; max_scene_cl = S; }
    cudaStream_t st = (cudaStream_t)stream;
    { unlabeled_argmin_kernel<<<sgb_div_up(S, 128), 128, 0, st>>>(dist, row_off, nbr, eid, S, amin_ws, scene_cl_off, n_scenes); SGB_COUNT_LAUNCH(); }
    const size_t state_bytes = ((size_t)6 * max_scene_segs + 3 * (size_t)max_scene_cl) * sizeof(int);
    const int in_smem = state_bytes <= 200 * 1024;
    if (in_smem && state_bytes > 40 * 1024)
        SGB_OPT_IN_SMEM(unlabeled_union_kernel);
    { unlabeled_union_kernel<<<n_scenes, UU_THREADS, in_smem ? state_bytes : 0, st>>>(amin_ws, S, roots_cur, uf, S1, in_smem, amin_ws + S,
                                                                                      scene_seg_off, scene_cl_off); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// phase B of group_unlabeled_clusters: unl [n_unl] dense ids of the unlabeled clusters (ascending), cand [n_unl,S] all cluster
// ids sorted by increasing cloud distance from the respective unlabeled cluster
extern "C" int sgb_group_unlabeled_phase_b(const int* unl, int n_unl, const int* cand, int S, const int* roots_cur, int* uf, int S1,
                                           void* stream) {
    if (n_unl < 0 || S <= 0 || S1 <= 0 || !roots_cur || !uf) return SGB_ERR_INVALID;
    if (n_unl == 0) return SGB_OK;
    if (!unl || !cand) return SGB_ERR_INVALID;
    { unlabeled_phase_b_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(unl, n_unl, cand, S, roots_cur, uf, S1); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

// phase B candidate ranking (see phase_b_rank_kernel).  cloud_idx [S,P]: sampled point ids of every cluster (sgb_cluster_cloud_indices),
// unl [n_unl] ascending dense ids of the unlabeled clusters, cand [n_unl,width] (width >= clusters of the largest scene, <= 4096).
extern "C" int sgb_phase_b_rank(const float* xyz, int stride, const int* cloud_idx, int P, const int* unl, int n_unl,
                                const int* scene_cl_off, int n_scenes, int S, int* cand, int width, void* stream) {
    if (n_unl < 0 || S <= 0 || P <= 0 || stride < 3 || width <= 0 || n_scenes < 1) return SGB_ERR_INVALID;
    if (n_unl == 0) return SGB_OK;
    if (!xyz || !cloud_idx || !unl || !cand) return SGB_ERR_INVALID;
    if (width > PB_MAX) return SGB_ERR_UNSUPPORTED;
    { phase_b_rank_kernel<<<n_unl, PB_THREADS, 0, (cudaStream_t)stream>>>(xyz, stride, cloud_idx, P, unl, n_scenes > 1 ? scene_cl_off : nullptr,
                                                                          n_scenes, S, cand, width); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}

extern "C" int sgb_export_labels(const long long* unmap, int n_raw, const int* seg_of_point, const int* seg2cl, const int* cl_rootpt,
                                 const int* cl_ins, const int* cl_sem, int* out_seg, int* out_ins, int* out_sem, void* stream) {
    return sgb_export_labels_scenes(unmap, n_raw, seg_of_point, seg2cl, cl_rootpt, cl_ins, cl_sem, out_seg, out_ins, out_sem, nullptr, 1, stream);
}
// scene batch: scene_pt_off [n_scenes+1] (device) point range of every scene; segment labels are root point ids inside the scene
extern "C" int sgb_export_labels_scenes(const long long* unmap, int n_raw, const int* seg_of_point, const int* seg2cl, const int* cl_rootpt,
                                        const int* cl_ins, const int* cl_sem, int* out_seg, int* out_ins, int* out_sem,
                                        const int* scene_pt_off, int n_scenes, void* stream) {
    if (n_raw <= 0 || !seg_of_point || !seg2cl || !cl_rootpt || !cl_ins || !cl_sem || n_scenes < 1) return SGB_ERR_INVALID;
    { export_labels_kernel<<<sgb_div_up(sgb_div_up(n_raw, EXP_V), 256), 256, 0, (cudaStream_t)stream>>>(unmap, n_raw, seg_of_point, seg2cl, cl_rootpt,
                                                                                   cl_ins, cl_sem, out_seg, out_ins, out_sem,
                                                                                   n_scenes > 1 ? scene_pt_off : nullptr, n_scenes); SGB_COUNT_LAUNCH(); }
    SGB_CHECK_LAUNCH();
    return SGB_OK;
}
