// TEST INFRASTRUCTURE — extern "C" driver around the UNMODIFIED reference C++ cores
// (/root/reference/kpconv/tf_custom_ops/**/{neighbors,grid_subsampling}.cpp, cpp_wrappers/**/grid_subsampling.cpp,
// cpp_utils/cloud/cloud.cpp, vendored nanoflann 1.3.0).  Built by oracle/build_ref.py into oracle/_ref/ from the
// sources where they lie; nothing here is shipped in the product.  This file only marshals flat arrays into
// std::vector<PointXYZ> the way tf_batch_neighbors.cpp:40-116, tf_batch_subsampling.cpp:30-122 and
// wrapper.cpp:92-276 do.
#include <cstring>
#include <vector>
#include "tf_custom_ops/tf_neighbors/neighbors/neighbors.h"

// the tf variant (points only / single label) and the cpp_wrappers variant (multi-column labels, verbose) share
// the function name with different signatures -> both can be linked (C++ overloads).
void grid_subsampling(std::vector<PointXYZ>&, std::vector<PointXYZ>&, std::vector<float>&, std::vector<float>&,
                      std::vector<int>&, std::vector<int>&, float);
void batch_grid_subsampling(std::vector<PointXYZ>&, std::vector<PointXYZ>&, std::vector<float>&, std::vector<float>&,
                            std::vector<int>&, std::vector<int>&, std::vector<int>&, std::vector<int>&, float);
void grid_subsampling(std::vector<PointXYZ>&, std::vector<PointXYZ>&, std::vector<float>&, std::vector<float>&,
                      std::vector<int>&, std::vector<int>&, float, int);

static std::vector<PointXYZ> to_points(const float* xyz, int n) {
    std::vector<PointXYZ> v((size_t)n);
    if (n) std::memcpy((void*)v.data(), xyz, (size_t)n * 3 * sizeof(float));
    return v;
}

static std::vector<int> g_nb;
static std::vector<PointXYZ> g_sub;
static std::vector<float> g_subf;
static std::vector<int> g_subc, g_subb;

extern "C" {
// returns the row width W; rows via ref_neighbors_copy
int ref_batch_neighbors(const float* q, int nq, const float* s, int ns, const int* qb, const int* sb, int nb, float radius, int nanoflann) {
    std::vector<PointXYZ> Q = to_points(q, nq), S = to_points(s, ns);
    std::vector<int> QB(qb, qb + nb), SB(sb, sb + nb);
    g_nb.clear();
    if (nanoflann) batch_nanoflann_neighbors(Q, S, QB, SB, g_nb, radius);
    else batch_ordered_neighbors(Q, S, QB, SB, g_nb, radius);
    return nq ? (int)(g_nb.size() / (size_t)nq) : 0;
}
int ref_ordered_neighbors(const float* q, int nq, const float* s, int ns, float radius) {
    std::vector<PointXYZ> Q = to_points(q, nq), S = to_points(s, ns);
    g_nb.clear();
    ordered_neighbors(Q, S, g_nb, radius);
    return nq ? (int)(g_nb.size() / (size_t)nq) : 0;
}
void ref_neighbors_copy(int* out) { if (!g_nb.empty()) std::memcpy(out, g_nb.data(), g_nb.size() * sizeof(int)); }

// tf variant: points only, batched.  returns M; out_batches[nb] filled
int ref_batch_grid_subsampling(const float* xyz, int n, const int* batches, int nb, float dl, int* out_batches) {
    std::vector<PointXYZ> P = to_points(xyz, n);
    std::vector<float> f, sf;
    std::vector<int> c, sc, B(batches, batches + nb);
    g_sub.clear(); g_subb.clear();
    batch_grid_subsampling(P, g_sub, f, sf, c, sc, B, g_subb, dl);
    for (int i = 0; i < nb; ++i) out_batches[i] = g_subb[i];
    return (int)g_sub.size();
}
// cpp_wrappers variant: optional features [n,fdim] and labels [n,ldim]
int ref_grid_subsampling(const float* xyz, int n, const float* feat, int fdim, const int* cls, int ldim, float dl) {
    std::vector<PointXYZ> P = to_points(xyz, n);
    std::vector<float> f; if (feat && fdim) f.assign(feat, feat + (size_t)n * fdim);
    std::vector<int> c; if (cls && ldim) c.assign(cls, cls + (size_t)n * ldim);
    g_sub.clear(); g_subf.clear(); g_subc.clear();
    grid_subsampling(P, g_sub, f, g_subf, c, g_subc, dl, 0);
    return (int)g_sub.size();
}
void ref_subsampling_copy(float* xyz, float* feat, int* cls) {
    if (xyz && !g_sub.empty()) std::memcpy(xyz, (void*)g_sub.data(), g_sub.size() * 3 * sizeof(float));
    if (feat && !g_subf.empty()) std::memcpy(feat, g_subf.data(), g_subf.size() * sizeof(float));
    if (cls && !g_subc.empty()) std::memcpy(cls, g_subc.data(), g_subc.size() * sizeof(int));
}
}
