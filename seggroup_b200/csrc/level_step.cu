// One clustering level of the segment graph in ONE library call (host-side composition of the graph.cu entry points):
//   [group_nearby | group_unlabeled_step | nothing] -> level_build -> level_children -> update_adj -> sym_csr
// i.e. seggroup/model.py:752-770 (and 802-815, 843-856, 441-470) between "distances are known" and "features can be
// pooled to the new level".  The Python layer used to drive these ~45 kernels through five entry points and two
// blocking read-backs per level; here the two read-backs (cluster count, edge count: they size the next launches) are
// the only host synchronisation and the launches are issued from C, which is what makes the narrow graph kernels
// cheap to drive from several scene threads at once.
#include "common.cuh"

extern "C" size_t sgb_level_step_ws_bytes(int S1, int S_old) {
    const int s = S_old > S1 ? S_old : S1;
    size_t a = sgb_level_ws_bytes(S1);
    const size_t b = sgb_children_ws_bytes(s), c = sgb_update_adj_ws_bytes(s), d = sgb_sym_csr_ws_bytes(s);
    if (b > a) a = b;
    if (c > a) a = c;
    if (d > a) a = d;
    return a + 256;
}


// mode 0: group_nearby(adj_old, dist, th); 1: group_unlabeled_step(dist, csr_old); 2: no grouping (level of the input graph).
// edges/E/map: the edge list to re-map for the new adjacency; map == NULL -> old2new of this step (edges = adj_old).
// roots_old == NULL (first level): no children CSR.  counts_host [4] <- S_new, A_new, #unlabeled clusters, status word.
// All cluster-sized outputs are sized for S1 (offsets S1 + 1), order for N, adj_new for E rows, csr_nbr / csr_eid for 2 E.
extern "C" int sgb_level_step(int mode, const int* adj_old, int A_old, const int* roots_old, int S_old, const float* dist, float th,
                              int sweep_cap, const int* csr_off_old, const int* csr_nbr_old, const int* csr_eid_old,
                              const int* edges, int E, const int* map,
                              int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                              int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                              int* cl_ins, int* cl_sem, int* cl_rootpt, int* old2new, int* child_off, int* child_list,
                              int* adj_new, int* csr_off, int* csr_nbr, int* csr_eid,
                              int* status, int* counts_dev, int* counts_host, void* ws, size_t ws_bytes, void* stream) {
    return sgb_level_step_scenes(mode, adj_old, A_old, roots_old, S_old, dist, th, sweep_cap, csr_off_old, csr_nbr_old, csr_eid_old,
                                 edges, E, map, uf, S1, N, seg_off, seg_members, seg_of_pos, roots, seg2cl, cl_seg_off, cl_seg_list,
                                 cl_pt_off, order, cl_ins, cl_sem, cl_rootpt, old2new, child_off, child_list, adj_new, csr_off, csr_nbr,
                                 csr_eid, status, counts_dev, counts_host, nullptr, nullptr, nullptr, 1, S1, S_old, ws, ws_bytes, stream);
}

// Scene batch (several scenes concatenated into one block-diagonal graph, ids offset): scene_seg_off [n_scenes+1] (device) level-1
// segment range of every scene, scene_cl_off_old [n_scenes+1] (device) its cluster range at the OLD level, scene_cl_off_new
// [n_scenes+1] (device, output) at the new one; max_scene_segs / max_scene_cl_old size the per-scene shared-memory state of the
// order-dependent replays, which run one CTA per scene.  counts_dev / counts_host need 4 + n_scenes + 1 ints: the new cluster
// offsets of the scenes follow the four counters.  n_scenes == 1 with NULL scene arrays == sgb_level_step.
extern "C" int sgb_level_step_scenes(int mode, const int* adj_old, int A_old, const int* roots_old, int S_old, const float* dist, float th,
                                     int sweep_cap, const int* csr_off_old, const int* csr_nbr_old, const int* csr_eid_old,
                                     const int* edges, int E, const int* map,
                                     int* uf, int S1, int N, const int* seg_off, const int* seg_members, const int* seg_of_pos,
                                     int* roots, int* seg2cl, int* cl_seg_off, int* cl_seg_list, int* cl_pt_off, int* order,
                                     int* cl_ins, int* cl_sem, int* cl_rootpt, int* old2new, int* child_off, int* child_list,
                                     int* adj_new, int* csr_off, int* csr_nbr, int* csr_eid,
                                     int* status, int* counts_dev, int* counts_host,
                                     const int* scene_seg_off, const int* scene_cl_off_old, int* scene_cl_off_new, int n_scenes,
                                     int max_scene_segs, int max_scene_cl_old, void* ws, size_t ws_bytes, void* stream) {
    if (S1 <= 0 || N <= 0 || !uf || !status || !counts_dev || !counts_host || !ws || n_scenes < 1) return SGB_ERR_INVALID;
    if (n_scenes > 1 && (!scene_seg_off || !scene_cl_off_new || (mode != 2 && !scene_cl_off_old))) return SGB_ERR_INVALID;
    if (ws_bytes < sgb_level_step_ws_bytes(S1, S_old)) return SGB_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const bool batch = scene_seg_off != nullptr;
    int rc;
    if (mode == 0) {
        if ((rc = sgb_group_nearby_scenes(adj_old, A_old, roots_old, S_old, dist, th, uf, S1, sweep_cap, status, scene_seg_off,
                                          scene_cl_off_old, n_scenes, max_scene_segs, stream))) return rc;
    } else if (mode == 1) {
        // amin scratch: the head of ws (2 * S_old ints <= 6 * (S1 + 1)); the level workspace is used afterwards, stream order keeps them apart
        if ((rc = sgb_group_unlabeled_step_scenes(dist, csr_off_old, csr_nbr_old, csr_eid_old, S_old, roots_old, uf, S1, (int*)ws,
                                                  scene_seg_off, scene_cl_off_old, n_scenes, max_scene_segs, max_scene_cl_old, stream))) return rc;
    }
    if ((rc = sgb_level_build_scenes(uf, S1, N, seg_off, seg_members, seg_of_pos, roots, seg2cl, cl_seg_off, cl_seg_list, cl_pt_off, order,
                                     cl_ins, cl_sem, cl_rootpt, counts_dev, scene_seg_off, n_scenes, batch ? scene_cl_off_new : nullptr,
                                     ws, ws_bytes, stream))) return rc;
    // sgb_level_build leaves counts_dev[0] = clusters and counts_dev[2] = clusters without a label
    if (batch && scene_cl_off_new == counts_dev + 4) {         // the caller laid the scene offsets out behind the counters: one copy
        SGB_CUDA(cudaMemcpyAsync(counts_host, counts_dev, (size_t)(4 + n_scenes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    } else {
        SGB_CUDA(cudaMemcpyAsync(counts_host, counts_dev, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (batch) SGB_CUDA(cudaMemcpyAsync(counts_host + 4, scene_cl_off_new, (size_t)(n_scenes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    SGB_CUDA(cudaStreamSynchronize(st));
    const int S_new = counts_host[0];
    if (S_new <= 0) return SGB_ERR_INVALID;
    if (roots_old) {
        if ((rc = sgb_level_children(roots_old, S_old, seg2cl, S_new, old2new, child_off, child_list, ws, ws_bytes, stream))) return rc;
    }
    const int* m = map ? map : old2new;
    if ((rc = sgb_update_adj(edges, E, m, S_new, adj_new, counts_dev, ws, ws_bytes, stream))) return rc;
    SGB_CUDA(cudaMemcpyAsync(counts_host + 1, counts_dev + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGB_CUDA(cudaMemcpyAsync(counts_host + 3, status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SGB_CUDA(cudaStreamSynchronize(st));
    const int A_new = counts_host[1];
    if ((rc = sgb_sym_csr(adj_new, A_new, S_new, csr_off, csr_nbr, csr_eid, ws, ws_bytes, stream))) return rc;
    return SGB_OK;
}
