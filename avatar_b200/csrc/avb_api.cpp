// avb_api.cpp -- host side of the C ABI declared in include/avatar_b200.h.
//
// Owns the model preparation (the data layout the kernels read), device buffers, the per-call
// kernel schedule of AvatarOptimizer::optimize (AvatarOptimizer.cpp:1246-1517) and the transfers.
// There is no CPU compute path here: every fit entry point enqueues the CUDA kernels of
// avb_kernels.cu and fails with AVB_ERR_CUDA if that is impossible.
#include "../../include/avatar_b200.h"
#include "avb_kernels.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace avb;

static_assert(sizeof(avb_stats) == sizeof(FrameStats), "avb_stats must mirror FrameStats");
static_assert(AVB_MAX_ASSIGN == AVB_MAX_ASSIGN_, "assign width");
static_assert(AVB_MAX_JOINTS == kMaxJ && AVB_MAX_SHAPE_KEYS == kMaxK, "limits");

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(AVB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// dense lower Cholesky, row-major; false if not positive definite
bool chol_lower(const std::vector<double>& A, int n, std::vector<double>& L) {
    L.assign((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        L[(size_t)j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
            L[(size_t)i * n + j] = s / d;
        }
    }
    return true;
}

template <class T>
cudaError_t dev_upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dst), bytes);
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}
}  // namespace

namespace {
// NCCL is loaded at run time (libnccl.so.2: the system library, or the one a host framework already mapped), so that
// single-GPU users of libavatar_b200.so need no NCCL at all.
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
            api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
        }
    }
    return api;
}
int nccl_fail(const char* what, int rc) {
    return fail(AVB_ERR_CUDA, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error"));
}
}  // namespace

/* ------------------------------------------------------------------------------------------ */
struct avb_model {
    int V = 0, J = 0, K = 0, F = 0, P = 0, nx = 0, max_depth = 0;
    std::vector<double> vt, sk_w, jbase, jreg, Sp;
    std::vector<float> sd;
    std::vector<uint8_t> sk_j, sk_n;
    std::vector<uint32_t> anc_mask;
    std::vector<int> parent, depth, faces;
    std::vector<int> main_joint;  // assignedJoints[v][0].second
    int gmmC = 0, gmmD = 0;
    std::vector<double> gmm_mean, gmm_prec, gmm_prec_cho, gmm_clog;
};

struct avb_fitter {
    const avb_model* model = nullptr;
    int device = 0, max_batch = 0, num_sms = 148;
    int64_t max_points = 0;
    int num_parts = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaEvent_t hx_ev = nullptr; bool hx_busy = false;   // recorded after the H2D copy that reads the h_x staging buffer
    DevModel dm{};
    DevParts dp{};
    std::vector<void*> allocs;  // everything cudaMalloc'ed, freed in destroy
    std::vector<void*> pinned;
    // per-batch device buffers
    double* d_data = nullptr; int* d_labels = nullptr;
    double* d_x = nullptr; double* d_xdbg = nullptr;
    double* d_cloud = nullptr; double* d_jpos = nullptr; double* d_jtrans = nullptr;
    uint8_t* d_vis = nullptr; int* d_pv_idx = nullptr; double* d_pv_xyz = nullptr; int* d_pv_start = nullptr;
    long long pv_stride = 0;
    int* d_nn = nullptr; int* d_cnt = nullptr; unsigned long long* d_sum = nullptr; double* d_qpart = nullptr;
    int* d_range = nullptr; double* d_Hcur = nullptr; FrameStats* d_stats = nullptr;
    double *d_xt = nullptr, *d_tab = nullptr, *d_part = nullptr, *d_cpart = nullptr, *d_gcur = nullptr;
    unsigned short* d_mlist = nullptr; int4* d_chunks = nullptr; int2* d_gruns = nullptr; LmState* d_state = nullptr;
    float* d_rec = nullptr; int* d_gstart = nullptr; int maxrb = 0, rec_stride = 0, rec_rs = 0;
    int nn_stage_cap = 0;
    int nn_f32 = 1;                // nn_kernel: fp32 pre-scan with exact fp64 fallback (AVB_NN_F32=0: plain fp64 scan)
    float4* d_pv_f32 = nullptr; float* d_pv_rmax = nullptr;
    int fused64 = 0;               // fp64 flow path, AVB_FUSED=1: fused record + Gram tasks (no d_rec round trip).  Measured NOT
                                   // faster (60.5 k vs 61.2 k frames/s): a chunk fills 75 % of the CTA's threads, DESIGN.md 5.4
    float* d_data_f32 = nullptr;   // avb_upload_batch_f32: the cloud as uploaded floats
    bool data_is_f32 = false;      // the resident cloud lives in d_data_f32 (nn_kernel widens on load); d_data is filled on demand
    int maxc = 0, tabD = 0, max_nj = 0, chunk_verts = 192, chunk_verts_tc = 256;   // vertices per Gram chunk: fp64 path (3 CTAs/SM) / tensor path
    long long pstride = 0;
    // lm_flow_kernel work queue
    unsigned long long* d_qslots = nullptr; unsigned int* d_qctrl = nullptr; int *d_rows_left = nullptr, *d_gram_left = nullptr; double* d_prior_out = nullptr; int prior_stride = 0; int prior_task_max = 32;
    unsigned long long* d_qprof = nullptr; unsigned int qcap = 0; bool use_flow = true;
    int flow_occ_default[2] = {3, 2};   // CTAs per SM of the flow kernel: [fp64 path, tensor path] (measured best)
    unsigned int* h_qctrl = nullptr;
    std::vector<int> group_nj, group_nv;   // Jacobian column groups: joints and model vertices per group
    int *d_chunk_frame = nullptr, *d_chunk_count = nullptr, *d_chunk_qblock = nullptr, *d_frame_qblock = nullptr;
    long long* d_chunk_begin = nullptr;
    double *d_dump_cost = nullptr, *d_dump_grad = nullptr, *d_dump_H = nullptr;
    // pinned host staging
    double* h_x = nullptr; FrameStats* h_stats = nullptr;
    int *h_chunk_frame = nullptr, *h_chunk_count = nullptr, *h_chunk_qblock = nullptr, *h_frame_qblock = nullptr;
    long long* h_chunk_begin = nullptr;
    int max_chunks = 0;
    int64_t max_qblocks = 0;
    // cloud construction from images (avb_upload_depth_batch): staging grown on demand
    float* d_depth = nullptr; uint8_t* d_parts = nullptr; int* d_roi = nullptr; int* d_strip_count = nullptr;
    long long* d_strip_offset = nullptr; int* d_bad_label = nullptr;
    size_t img_cap = 0, strip_cap = 0;
    RTreeNode* d_rt_nodes = nullptr; uint8_t* d_rt_leaf = nullptr; int rt_nodes = 0, rt_leaves = 0, rt_parts = 0;
    // RTree::postProcess scratch (grown on demand) and the per-frame centre-of-mass state of the chained pipeline
    int *d_pp_arena = nullptr, *d_pp_stack = nullptr, *d_pp_ovf = nullptr; double* d_pp_com = nullptr; size_t pp_cap = 0; int pp_parts = 0;
    const uint8_t* d_vpart = nullptr; float* d_proj = nullptr; int* d_order = nullptr; unsigned* d_win = nullptr;
    float* d_rdepth = nullptr; uint8_t* d_rparts = nullptr; int* d_rfaces = nullptr; size_t render_cap = 0;
    const int *d_vf_start = nullptr, *d_vf_list = nullptr; int max_valence = 0;   // faces incident to every vertex (renderLambert)
    int* d_rank_of = nullptr; float* d_vlam = nullptr; unsigned* d_win_l = nullptr; uint8_t* d_rlam = nullptr; size_t lam_cap = 0;
    cudaEvent_t nev[4] = {};   // around the three renderer kernels
    cudaEvent_t rev[2] = {};   // around the RTree kernels of the last prediction
    cudaEvent_t cev[4] = {};   // around cloud_count_kernel and cloud_compact_kernel of the last avb_upload_depth_batch
    std::vector<int> h_strip_count; std::vector<long long> h_strip_offset;
    // state of the uploaded batch
    int batch = 0, num_chunks = 0;
    int64_t total_points = 0;
    std::vector<int64_t> offsets;
    int launches = 0;
    int n_events = 0;
    // per-kernel profiling (avb_set_profiling): events around every launch of the next avb_fit_resident
    bool profile = false;
    std::vector<cudaEvent_t> pev;   // pairs (begin, end)
    std::vector<int> pcls;          // kernel class of each pair
    // tracking mode (avb_track_sequence)
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> copied;
    double* d_xseq = nullptr; FrameStats* d_stats_seq = nullptr; int seq_cap = 0;
    FrameStats* h_stats_seq = nullptr; double* h_xseq = nullptr;
    int last_icp = 0;
    // multi-GPU: NCCL communicator of this fitter (avb_fitter_comm_init) and the all-gather buffers
    void* nccl_comm = nullptr; int comm_rank = 0, comm_size = 1;
    double* d_gather = nullptr; double* h_gather = nullptr; double* d_send = nullptr;
    cudaStream_t gather_stream = nullptr; cudaEvent_t ev_send = nullptr, ev_gdone = nullptr; bool gather_pending = false;
};

namespace {

template <class T>
int dev_alloc(avb_fitter* ft, T** p, size_t count) {
    *p = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(AVB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    ft->allocs.push_back(*p);
    return AVB_OK;
}
template <class T>
int pin_alloc(avb_fitter* ft, T** p, size_t count) {
    *p = nullptr;
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T), cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(AVB_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    ft->pinned.push_back(*p);
    return AVB_OK;
}
template <class T>
void dev_release(avb_fitter* ft, T** p) {   // frees one dev_alloc'ed buffer now (not at destroy)
    if (!*p) return;
    auto it = std::find(ft->allocs.begin(), ft->allocs.end(), static_cast<void*>(*p));
    if (it != ft->allocs.end()) ft->allocs.erase(it);
    cudaFree(*p);
    *p = nullptr;
}
template <class T>
void pin_release(avb_fitter* ft, T** p) {
    if (!*p) return;
    auto it = std::find(ft->pinned.begin(), ft->pinned.end(), static_cast<void*>(*p));
    if (it != ft->pinned.end()) ft->pinned.erase(it);
    cudaFreeHost(*p);
    *p = nullptr;
}
// NN chunk tables (device + pinned mirrors) for at least `need` chunks; the stream must be idle when they grow
int ensure_chunk_tables(avb_fitter* ft, size_t need) {
    if (need <= (size_t)ft->max_chunks) return AVB_OK;
    cudaError_t e = cudaStreamSynchronize(ft->stream);
    if (e != cudaSuccess) return fail(AVB_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    dev_release(ft, &ft->d_chunk_frame); dev_release(ft, &ft->d_chunk_count); dev_release(ft, &ft->d_chunk_qblock);
    dev_release(ft, &ft->d_chunk_begin);
    pin_release(ft, &ft->h_chunk_frame); pin_release(ft, &ft->h_chunk_count); pin_release(ft, &ft->h_chunk_qblock);
    pin_release(ft, &ft->h_chunk_begin);
    ft->max_chunks = 0;
    need += need / 4 + 64;
    int rc = dev_alloc(ft, &ft->d_chunk_frame, need);
    if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_chunk_count, need);
    if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_chunk_qblock, need);
    if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_chunk_begin, need);
    if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_chunk_frame, need);
    if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_chunk_count, need);
    if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_chunk_qblock, need);
    if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_chunk_begin, need);
    if (rc != AVB_OK) return rc;
    ft->max_chunks = (int)need;
    return AVB_OK;
}
template <class T>
int dev_put(avb_fitter* ft, const T** p, const std::vector<T>& src) {
    T* d = nullptr;
    cudaError_t e = dev_upload(&d, src);
    if (d) ft->allocs.push_back(d);
    if (e != cudaSuccess) return fail(AVB_ERR_CUDA, std::string("model upload: ") + cudaGetErrorString(e));
    *p = d;
    return AVB_OK;
}

// Static column-group schedule.  Vertices whose skinning joints share the same ancestor set get
// the same sparse Jacobian column pattern; groups are merged greedily until at most `max_groups`
// remain, minimising sum_g |g| * Lp(g)^2 (the dense-within-group J^T J work).
void build_groups(const avb_model& m, int max_groups, std::vector<int>& gorder, std::vector<int>& gvstart,
                  std::vector<int>& gjoints, std::vector<int>& gnj) {
    const int V = m.V, J = m.J, K = m.K;
    std::vector<uint32_t> sig(V);
    for (int v = 0; v < V; ++v) {
        uint32_t s = 0;
        for (int q = 0; q < m.sk_n[v]; ++q) s |= m.anc_mask[m.sk_j[4 * (size_t)v + q]];
        sig[v] = s;
    }
    struct Grp { uint32_t mask; long long n; };
    std::vector<Grp> groups;
    {
        std::vector<uint32_t> u(sig);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        for (uint32_t s : u) groups.push_back({s, 0});
        for (int v = 0; v < V; ++v)
            for (auto& g : groups)
                if (g.mask == sig[v]) { ++g.n; break; }
    }
    auto lp = [&](uint32_t mask) {   // padded record length of a group (the Gram matrix is lp x lp)
        const int L = 3 * __builtin_popcount(mask) + 3 * K + 7;
        return (L + 7) & ~7;
    };
    auto cost = [&](uint32_t mask, long long n) { return (double)n * lp(mask) * lp(mask); };
    while ((int)groups.size() > max_groups) {
        double best = 1e300;
        int bi = 0, bj = 1;
        for (size_t i = 0; i < groups.size(); ++i)
            for (size_t j = i + 1; j < groups.size(); ++j) {
                const uint32_t mm = groups[i].mask | groups[j].mask;
                const double inc = cost(mm, groups[i].n + groups[j].n) - cost(groups[i].mask, groups[i].n) -
                                   cost(groups[j].mask, groups[j].n);
                if (inc < best) { best = inc; bi = (int)i; bj = (int)j; }
            }
        groups[bi].mask |= groups[bj].mask;
        groups[bi].n += groups[bj].n;
        groups.erase(groups.begin() + bj);
        // absorb groups that became subsets for free
        for (size_t j = 0; j < groups.size();) {
            if ((int)j != bi && (groups[j].mask | groups[bi].mask) == groups[bi].mask &&
                lp(groups[j].mask) == lp(groups[bi].mask)) {
                groups[bi].n += groups[j].n;
                if ((int)j < bi) --bi;
                groups.erase(groups.begin() + j);
            } else {
                ++j;
            }
        }
    }
    std::sort(groups.begin(), groups.end(), [&](const Grp& a, const Grp& b) {
        const int pa = __builtin_popcount(a.mask), pb = __builtin_popcount(b.mask);
        return pa != pb ? pa < pb : a.mask < b.mask;
    });
    const int G = (int)groups.size();
    // assign every vertex to the cheapest group that covers its signature
    std::vector<int> vg(V, -1);
    for (int v = 0; v < V; ++v) {
        int best = -1;
        for (int g = 0; g < G; ++g)
            if ((sig[v] | groups[g].mask) == groups[g].mask && (best < 0 || lp(groups[g].mask) < lp(groups[best].mask))) best = g;
        vg[v] = best;
    }
    gorder.clear();
    gvstart.assign(G + 1, 0);
    for (int g = 0; g < G; ++g) {
        gvstart[g] = (int)gorder.size();
        for (int v = 0; v < V; ++v)
            if (vg[v] == g) gorder.push_back(v);
    }
    gvstart[G] = (int)gorder.size();
    gjoints.assign((size_t)G * kMaxJ, 0);
    gnj.assign(G, 0);
    for (int g = 0; g < G; ++g)
        for (int j = 0; j < J; ++j)
            if ((groups[g].mask >> j) & 1u) gjoints[(size_t)g * kMaxJ + gnj[g]++] = j;
}

}  // namespace

extern "C" {

const char* avb_last_error(void) { return g_err.c_str(); }

void avb_default_options(avb_options* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->icp_iters = 1;            // AvatarOptimizer.h:19
    o->max_iters_per_icp = 10;   // AvatarOptimizer.h:36
    o->beta_pose = 0.1;          // AvatarOptimizer.h:27
    o->beta_shape = 1.0;
    o->enable_occlusion = 1;     // AvatarOptimizer.h:39
    o->nn_step = 20;             // AvatarOptimizer.h:33 (unused by the inverted NN mode)
    o->function_tolerance = 1e-4;  // AvatarOptimizer.cpp:1333
    o->solver = AVB_SOLVER_GN_LM;
    o->jtj_precision = AVB_JTJ_FP64;   // parity path; AVB_JTJ_BF16_TENSOR = tcgen05 J^T J (faster, fits within ~3e-4 of this path)
}

int avb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int avb_model_create(const avb_model_desc* d, avb_model** out) {
    if (!d || !out) return fail(AVB_ERR_INVALID, "null argument");
    *out = nullptr;
    const int V = d->num_points, J = d->num_joints, K = d->num_shape_keys, F = d->num_faces;
    if (V <= 0 || V > 65535 || J <= 0 || J > kMaxJ || K < 0 || K > kMaxK || F < 0)
        return fail(AVB_ERR_INVALID, "unsupported model dimensions (V<=65535, J<=32, K<=16)");
    if (!d->base_cloud || !d->parent || !d->assign_start || !d->assign_joint || !d->assign_weight ||
        !d->joint_shape_reg_base || (K > 0 && (!d->key_clouds || !d->joint_shape_reg)) || (F > 0 && !d->mesh))
        return fail(AVB_ERR_INVALID, "null model array");
    auto m = new avb_model;
    m->V = V; m->J = J; m->K = K; m->F = F;
    m->P = 3 + 3 * J + K;
    m->nx = 3 + 4 * J + K;
    m->vt.assign(d->base_cloud, d->base_cloud + 3 * (size_t)V);
    m->sd.resize(3 * (size_t)V * K);
    for (size_t i = 0; i < m->sd.size(); ++i) m->sd[i] = (float)d->key_clouds[i];
    m->jbase.assign(d->joint_shape_reg_base, d->joint_shape_reg_base + 3 * (size_t)J);
    m->jreg.assign(d->joint_shape_reg, d->joint_shape_reg + 3 * (size_t)J * K);
    m->parent.assign(d->parent, d->parent + J);
    m->faces.assign(d->mesh, d->mesh + 3 * (size_t)F);
    for (int i = 0; i < 3 * F; ++i)
        if (m->faces[i] < 0 || m->faces[i] >= V) { delete m; return fail(AVB_ERR_INVALID, "face index out of range"); }
    // kinematic tree: depth and ancestor-or-self masks
    m->depth.assign(J, 0);
    m->anc_mask.assign(J, 0);
    for (int j = 0; j < J; ++j) {
        int dep = 0;
        uint32_t mask = 0;
        for (int a = j; a != -1; a = m->parent[a]) {
            if (a < 0 || a >= J || dep > J) { delete m; return fail(AVB_ERR_INVALID, "bad kinematic tree"); }
            mask |= 1u << a;
            ++dep;
        }
        m->depth[j] = dep - 1;
        m->anc_mask[j] = mask;
        m->max_depth = std::max(m->max_depth, dep - 1);
    }
    // Sp[j] = S[j] - S[parent j], Sp[0] = 0 (AvatarOptimizer.cpp:240-243)
    m->Sp.assign((size_t)J * 3 * K, 0.0);
    for (int j = 0; j < J; ++j) {
        if (m->parent[j] < 0) continue;
        for (int e = 0; e < 3 * K; ++e)
            m->Sp[(size_t)j * 3 * K + e] = m->jreg[(size_t)3 * j * K + e] - m->jreg[(size_t)3 * m->parent[j] * K + e];
    }
    // skinning (assignedJoints, at most 4 per vertex)
    m->sk_w.assign(4 * (size_t)V, 0.0);
    m->sk_j.assign(4 * (size_t)V, 0);
    m->sk_n.assign(V, 0);
    m->main_joint.assign(V, 0);
    for (int v = 0; v < V; ++v) {
        const int s = d->assign_start[v], e = d->assign_start[v + 1];
        if (e - s < 1 || e - s > AVB_MAX_ASSIGN) {
            delete m;
            return fail(AVB_ERR_INVALID, "every vertex needs 1..4 assigned joints (MAX_ASSIGN, AvatarOptimizer.cpp:164)");
        }
        m->sk_n[v] = (uint8_t)(e - s);
        for (int q = s; q < e; ++q) {
            if (d->assign_joint[q] < 0 || d->assign_joint[q] >= J) { delete m; return fail(AVB_ERR_INVALID, "assigned joint out of range"); }
            m->sk_j[4 * (size_t)v + q - s] = (uint8_t)d->assign_joint[q];
            m->sk_w[4 * (size_t)v + q - s] = d->assign_weight[q];
        }
        m->main_joint[v] = d->assign_joint[s];
    }
    // GaussianMixture::load maths (GaussianMixture.cpp:22-76)
    if (d->gmm_components > 0) {
        const int C = d->gmm_components, D = d->gmm_dims;
        if (D != 3 * (J - 1) || !d->gmm_weight || !d->gmm_mean || !d->gmm_cov) {
            delete m;
            return fail(AVB_ERR_INVALID, "pose prior must have 3(J-1) dimensions");
        }
        m->gmmC = C; m->gmmD = D;
        m->gmm_mean.assign(d->gmm_mean, d->gmm_mean + (size_t)C * D);
        m->gmm_clog.assign(C, 0.0);
        m->gmm_prec.assign((size_t)C * D * D, 0.0);
        m->gmm_prec_cho.assign((size_t)C * D * D, 0.0);
        const double log_sqrt_2_pi_n = D * 0.5 * std::log(2 * M_PI);
        double minDet = 1.79769313486231570e308;
        std::vector<double> cov((size_t)D * D), L, Linv((size_t)D * D), inv((size_t)D * D), Lp;
        for (int c = 0; c < C; ++c) {
            std::copy(d->gmm_cov + (size_t)c * D * D, d->gmm_cov + (size_t)(c + 1) * D * D, cov.begin());
            if (!chol_lower(cov, D, L)) { delete m; return fail(AVB_ERR_NUMERIC, "pose prior covariance is not positive definite"); }
            // Sigma^-1 = L^-T L^-1
            std::fill(Linv.begin(), Linv.end(), 0.0);
            for (int col = 0; col < D; ++col)
                for (int i = col; i < D; ++i) {
                    double s = (i == col) ? 1.0 : 0.0;
                    for (int k = col; k < i; ++k) s -= L[(size_t)i * D + k] * Linv[(size_t)k * D + col];
                    Linv[(size_t)i * D + col] = s / L[(size_t)i * D + i];
                }
            for (int a = 0; a < D; ++a)
                for (int b = 0; b <= a; ++b) {
                    double s = 0;
                    for (int k = a; k < D; ++k) s += Linv[(size_t)k * D + a] * Linv[(size_t)k * D + b];
                    inv[(size_t)a * D + b] = inv[(size_t)b * D + a] = s;
                }
            if (!chol_lower(inv, D, Lp)) { delete m; return fail(AVB_ERR_NUMERIC, "pose prior precision is not positive definite"); }
            std::copy(Lp.begin(), Lp.end(), m->gmm_prec_cho.begin() + (size_t)c * D * D);
            // the matrix the reference's residual and Jacobian imply is prec_cho prec_cho^T
            for (int a = 0; a < D; ++a)
                for (int b = 0; b <= a; ++b) {
                    double s = 0;
                    for (int k = 0; k <= b; ++k) s += Lp[(size_t)a * D + k] * Lp[(size_t)b * D + k];
                    m->gmm_prec[((size_t)c * D + a) * D + b] = m->gmm_prec[((size_t)c * D + b) * D + a] = s;
                }
            double det = 1.0;
            for (int k = 0; k < D; ++k) det *= L[(size_t)k * D + k];
            minDet = std::min(minDet, det);
            m->gmm_clog[c] = std::log(d->gmm_weight[c]) - log_sqrt_2_pi_n - std::log(det);
        }
        for (int c = 0; c < C; ++c) m->gmm_clog[c] += std::log(minDet);
    }
    *out = m;
    return AVB_OK;
}

void avb_model_destroy(avb_model* m) { delete m; }

int avb_model_dims(const avb_model* m, int32_t* V, int32_t* J, int32_t* K, int32_t* F) {
    if (!m) return fail(AVB_ERR_INVALID, "null model");
    if (V) *V = m->V;
    if (J) *J = m->J;
    if (K) *K = m->K;
    if (F) *F = m->F;
    return AVB_OK;
}
int avb_param_dim(const avb_model* m) { return m ? m->nx : 0; }
int avb_tangent_dim(const avb_model* m) { return m ? m->P : 0; }

int avb_model_get_prior(const avb_model* m, double* prec_cho, double* consts_log) {
    if (!m || m->gmmC <= 0) return fail(AVB_ERR_PRIOR, "model has no pose prior");
    if (prec_cho) std::copy(m->gmm_prec_cho.begin(), m->gmm_prec_cho.end(), prec_cho);
    if (consts_log) std::copy(m->gmm_clog.begin(), m->gmm_clog.end(), consts_log);
    return AVB_OK;
}

/* Eigen conventions: AngleAxisd::fromRotationMatrix -> Quaterniond */
void avb_rotmat_to_quat(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    double x, y, z, w;
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (m[7] - m[5]) * t;
        y = (m[2] - m[6]) * t;
        z = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double qq[3];
        t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        w = (m[3 * k + j] - m[3 * j + k]) * t;
        qq[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        qq[k] = (m[3 * k + i] + m[3 * i + k]) * t;
        x = qq[0]; y = qq[1]; z = qq[2];
    }
    // quaternion -> angle-axis (angle in [0, pi]) -> quaternion: canonical w >= 0
    double n = std::sqrt(x * x + y * y + z * z);
    if (n != 0) {
        const double angle = 2 * std::atan2(n, std::fabs(w));
        if (w < 0) n = -n;
        const double ha = 0.5 * angle, s = std::sin(ha);
        q[0] = s * (x / n); q[1] = s * (y / n); q[2] = s * (z / n); q[3] = std::cos(ha);
    } else {
        q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 1;
    }
}
void avb_quat_to_rotmat(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void* avb_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "cudaHostAlloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void avb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

void avb_fitter_destroy(avb_fitter* ft) {
    if (!ft) return;
    cudaSetDevice(ft->device);
    if (ft->stream) cudaStreamSynchronize(ft->stream);
    if (ft->gather_stream) { cudaStreamSynchronize(ft->gather_stream); cudaStreamDestroy(ft->gather_stream); }
    if (ft->ev_send) cudaEventDestroy(ft->ev_send);
    if (ft->ev_gdone) cudaEventDestroy(ft->ev_gdone);
    if (ft->nccl_comm && nccl().ok) nccl().CommDestroy(ft->nccl_comm);
    for (void* p : ft->allocs) cudaFree(p);
    for (void* p : ft->pinned) cudaFreeHost(p);
    for (auto& e : ft->ev)
        if (e) cudaEventDestroy(e);
    if (ft->hx_ev) cudaEventDestroy(ft->hx_ev);
    for (auto& e : ft->copied) cudaEventDestroy(e);
    for (auto& e : ft->pev) cudaEventDestroy(e);
    cudaFree(ft->d_depth); cudaFree(ft->d_parts); cudaFree(ft->d_roi); cudaFree(ft->d_strip_count);
    cudaFree(ft->d_strip_offset); cudaFree(ft->d_bad_label);
    for (auto& e : ft->cev) if (e) cudaEventDestroy(e);
    for (auto& e : ft->rev) if (e) cudaEventDestroy(e);
    cudaFree(ft->d_rt_nodes); cudaFree(ft->d_rt_leaf);
    cudaFree(ft->d_pp_arena); cudaFree(ft->d_pp_stack); cudaFree(ft->d_pp_ovf); cudaFree(ft->d_pp_com);
    cudaFree(ft->d_win); cudaFree(ft->d_rdepth); cudaFree(ft->d_rparts); cudaFree(ft->d_rfaces); cudaFree(ft->d_proj); cudaFree(ft->d_order);
    cudaFree(ft->d_rank_of); cudaFree(ft->d_vlam); cudaFree(ft->d_win_l); cudaFree(ft->d_rlam);
    for (auto& e : ft->nev) if (e) cudaEventDestroy(e);
    if (ft->copy_stream) cudaStreamDestroy(ft->copy_stream);
    if (ft->stream) cudaStreamDestroy(ft->stream);
    delete ft;
}

int avb_fitter_create(const avb_model* m, const avb_fitter_config* cfg, avb_fitter** out) {
    if (!m || !cfg || !out) return fail(AVB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->num_parts <= 0 || cfg->num_parts > kMaxParts || !cfg->part_map)
        return fail(AVB_ERR_INVALID, "num_parts must be in 1..64 and part_map non-null");
    if (cfg->max_batch <= 0 || cfg->max_total_points <= 0) return fail(AVB_ERR_INVALID, "capacities must be positive");
    for (int j = 0; j < m->J; ++j)
        if (cfg->part_map[j] < 0 || cfg->part_map[j] >= cfg->num_parts) return fail(AVB_ERR_INVALID, "part_map entry out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(AVB_ERR_CUDA, "no CUDA device available: avatar_b200 has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(AVB_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(AVB_ERR_CUDA, "avatar_b200 kernels are built for sm_100a only");
    {   // the kernels' shared-memory budgets, checked here instead of failing as an opaque launch error at fit time
        const size_t optin = (size_t)prop.sharedMemPerBlockOptin;
        if (pose_smem_bytes(m->V, m->J, m->K) > optin)
            return fail(AVB_ERR_INVALID, "model too large for pose_visibility_kernel (one visibility byte per vertex in shared memory)");
    }

    auto ft = new avb_fitter;
    ft->model = m;
    ft->device = cfg->device;
    ft->max_batch = cfg->max_batch;
    ft->max_points = cfg->max_total_points;
    ft->num_parts = cfg->num_parts;
    ft->num_sms = prop.multiProcessorCount;
    int rc = AVB_OK;
#define TRY(expr)                     \
    do {                              \
        rc = (expr);                  \
        if (rc != AVB_OK) {           \
            avb_fitter_destroy(ft);   \
            return rc;                \
        }                             \
    } while (0)
#define CUDA_TRY_FT(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            avb_fitter_destroy(ft);                                                               \
            return fail(AVB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
        }                                                                                         \
    } while (0)
    CUDA_TRY_FT(cudaStreamCreateWithFlags(&ft->stream, cudaStreamNonBlocking));
    for (auto& e : ft->ev) CUDA_TRY_FT(cudaEventCreate(&e));
    CUDA_TRY_FT(cudaEventCreateWithFlags(&ft->hx_ev, cudaEventDisableTiming));

    // ---- model -> device ----
    DevModel& dm = ft->dm;
    dm.V = m->V; dm.J = m->J; dm.K = m->K; dm.F = m->F; dm.P = m->P; dm.nx = m->nx; dm.max_depth = m->max_depth;
    TRY(dev_put(ft, &dm.vt, m->vt));
    TRY(dev_put(ft, &dm.sd, m->sd));
    {   // component-major copy for the thread-per-vertex sweep of pose_visibility_kernel (coalesced)
        std::vector<float> sdt(m->sd.size());
        const size_t V_ = (size_t)m->V, CK = (size_t)3 * m->K;
        for (size_t v = 0; v < V_; ++v)
            for (size_t q = 0; q < CK; ++q) sdt[q * V_ + v] = m->sd[v * CK + q];
        TRY(dev_put(ft, &dm.sdT, sdt));
    }
    TRY(dev_put(ft, &dm.sk_w, m->sk_w));
    TRY(dev_put(ft, &dm.sk_j, m->sk_j));
    TRY(dev_put(ft, &dm.sk_n, m->sk_n));
    TRY(dev_put(ft, &dm.anc_mask, m->anc_mask));
    TRY(dev_put(ft, &dm.parent, m->parent));
    TRY(dev_put(ft, &dm.depth, m->depth));
    {   // joints level by level (build_tables sweeps the kinematic tree one level per barrier)
        std::vector<int> lvl_start(m->max_depth + 2, 0), lvl_joint;
        for (int d = 0; d <= m->max_depth; ++d) {
            for (int j = 0; j < m->J; ++j)
                if (m->depth[j] == d) lvl_joint.push_back(j);
            lvl_start[d + 1] = (int)lvl_joint.size();
        }
        TRY(dev_put(ft, &dm.lvl_start, lvl_start));
        TRY(dev_put(ft, &dm.lvl_joint, lvl_joint));
    }
    TRY(dev_put(ft, &dm.jbase, m->jbase));
    TRY(dev_put(ft, &dm.jreg, m->jreg));
    TRY(dev_put(ft, &dm.Sp, m->Sp));
    TRY(dev_put(ft, &dm.faces, m->faces));
    dm.gmmC = m->gmmC; dm.gmmD = m->gmmD;
    TRY(dev_put(ft, &dm.gmm_mean, m->gmm_mean));
    TRY(dev_put(ft, &dm.gmm_prec, m->gmm_prec));
    TRY(dev_put(ft, &dm.gmm_clog, m->gmm_clog));
    {
        const size_t PT = (size_t)m->P * (m->P + 1) / 2;   // packed lower triangle, row major (the solve's storage)
        const int D = m->gmmD;
        std::vector<double> pfull((size_t)std::max(m->gmmC, 1) * PT, 0.0);
        for (int c = 0; c < m->gmmC; ++c)
            for (int r = 0; r < D; ++r)
                for (int q = 0; q <= r; ++q)
                    pfull[c * PT + (size_t)(6 + r) * (6 + r + 1) / 2 + 6 + q] = m->gmm_prec[((size_t)c * D + r) * D + q];
        TRY(dev_put(ft, &dm.gmm_pfull, pfull));
    }

    // ---- part tables (AvatarOptimizer.cpp:1223-1243) ----
    const int V = m->V, NP = cfg->num_parts;
    std::vector<int> part_start(NP + 1, 0), part_verts, first_part_at(V + 1, -1);
    for (int p = 0; p < NP; ++p) {
        part_start[p] = (int)part_verts.size();
        for (int v = 0; v < V; ++v)
            if (cfg->part_map[m->main_joint[v]] == p) part_verts.push_back(v);
    }
    part_start[NP] = (int)part_verts.size();
    for (int p = NP - 1; p >= 0; --p) first_part_at[part_start[p]] = p;
    {   // nn_kernel stages a prefix of the frame's visible model cloud in shared memory: as much as lets two CTAs share an
        // SM (about 4600 vertices; back-face culling leaves roughly half of the model visible), never more than V
        cudaDeviceProp pr2;
        CUDA_TRY_FT(cudaGetDeviceProperties(&pr2, cfg->device));
        const size_t per_cta = std::min<size_t>(pr2.sharedMemPerBlockOptin, (pr2.sharedMemPerMultiprocessor - 2 * 1024 - 2 * 24576) / 2);
        if (const char* e = std::getenv("AVB_NN_F32")) ft->nn_f32 = std::atoi(e) != 0;
        const size_t bpv = ft->nn_f32 ? 16 : 24;
        ft->nn_stage_cap = (int)std::min<size_t>((size_t)V, (per_cta - 256) / bpv);
        if (const char* e = std::getenv("AVB_NN_STAGE")) ft->nn_stage_cap = std::max(0, std::min(std::atoi(e), (int)((pr2.sharedMemPerBlockOptin - 256) / bpv)));
    }
    DevParts& dp = ft->dp;
    dp.numParts = NP;
    {   // part of every vertex: part_map[assignedJoints[v][0]] (the renderer's part mask, AvatarHelpers.cpp:162-169)
        std::vector<uint8_t> vpart((size_t)V);
        for (int v = 0; v < V; ++v) vpart[v] = (uint8_t)cfg->part_map[m->main_joint[v]];
        TRY(dev_put(ft, &ft->d_vpart, vpart));
    }
    {   // faces incident to every vertex (CSR, each face once per vertex, ascending): renderLambert's vertex normals
        std::vector<int> vs((size_t)V + 1, 0), vl;
        std::vector<std::vector<int>> inc((size_t)V);
        for (int t = 0; t < m->F; ++t)
            for (int c = 0; c < 3; ++c) {
                auto& l = inc[m->faces[3 * (size_t)t + c]];
                if (l.empty() || l.back() != t) l.push_back(t);
            }
        for (int v = 0; v < V; ++v) {
            vs[v] = (int)vl.size();
            vl.insert(vl.end(), inc[v].begin(), inc[v].end());
            ft->max_valence = std::max(ft->max_valence, (int)inc[v].size());
        }
        vs[V] = (int)vl.size();
        TRY(dev_put(ft, &ft->d_vf_start, vs));
        TRY(dev_put(ft, &ft->d_vf_list, vl));
    }
    TRY(dev_put(ft, &dp.part_start, part_start));
    TRY(dev_put(ft, &dp.part_verts, part_verts));
    TRY(dev_put(ft, &dp.first_part_at, first_part_at));
    int max_groups = 8;
    if (const char* e = std::getenv("AVB_GROUPS")) max_groups = std::max(1, std::min(kMaxGroups, std::atoi(e)));
    std::vector<int> gorder, gvstart, gjoints, gnj;
    build_groups(*m, max_groups, gorder, gvstart, gjoints, gnj);
    dp.numGroups = (int)gnj.size();
    ft->group_nj = gnj;
    for (size_t g = 0; g + 1 < gvstart.size(); ++g) ft->group_nv.push_back(gvstart[g + 1] - gvstart[g]);
    TRY(dev_put(ft, &dp.gorder, gorder));
    TRY(dev_put(ft, &dp.gvstart, gvstart));
    TRY(dev_put(ft, &dp.gjoints, gjoints));
    TRY(dev_put(ft, &dp.gnj, gnj));
    {   // scatter table of the partial reduction (lm_solve): group column -> tangent column [ p | 3 per joint | shape ]
        const int J = m->J, K = m->K, Pn = m->P, nTri = Pn * (Pn + 1) / 2;
        std::vector<int> gdoff(gnj.size() + 1, 0), gdest;
        for (size_t g = 0; g < gnj.size(); ++g) {
            const int nj = gnj[g], Lg = 3 + 3 * nj + K, nH = ((Lg + 1) >> 1) * (Lg + 1);   // tri_count(Lg), tri_decode order
            auto col = [&](int r) { return r < 3 ? r : (r < 3 + 3 * nj ? 3 + 3 * gjoints[g * kMaxJ + (r - 3) / 3] + (r - 3) % 3 : r + 3 * (J - nj)); };
            for (int idx = 0; idx < nH + Lg; ++idx) {
                int dest = -1;
                if (idx < nH) {
                    const int r = idx / (Lg + 1), c = idx - r * (Lg + 1);
                    int ra, rb;
                    bool ok = true;
                    if (c < Lg - r) { ra = r; rb = r + c; }
                    else { ra = Lg - 1 - r; rb = ra + (c - (Lg - r)); ok = ra != r; }
                    if (ok) {
                        const int ca = col(ra), cb = col(rb);
                        dest = ca >= cb ? ca * (ca + 1) / 2 + cb : cb * (cb + 1) / 2 + ca;
                    }
                } else {
                    dest = nTri + col(idx - nH);
                }
                gdest.push_back(dest);
            }
            gdoff[g + 1] = (int)gdest.size();
        }
        TRY(dev_put(ft, &dp.gdoff, gdoff));
        TRY(dev_put(ft, &dp.gdest, gdest));
    }

    // ---- batch buffers ----
    const size_t B = (size_t)cfg->max_batch, NT = (size_t)cfg->max_total_points;
    const size_t P = m->P, nx = m->nx, J = m->J;
    ft->pv_stride = (3 * (long long)V + 2 + 1) & ~1LL;
    TRY(dev_alloc(ft, &ft->d_data, 3 * NT));
    TRY(dev_alloc(ft, &ft->d_labels, NT));
    TRY(dev_alloc(ft, &ft->d_x, B * nx));
    TRY(dev_alloc(ft, &ft->d_xdbg, B * nx));
    TRY(dev_alloc(ft, &ft->d_cloud, B * 3 * V));
    TRY(dev_alloc(ft, &ft->d_jpos, B * 3 * J));
    TRY(dev_alloc(ft, &ft->d_jtrans, B * 12 * J));
    TRY(dev_alloc(ft, &ft->d_vis, B * V));
    TRY(dev_alloc(ft, &ft->d_pv_idx, B * V));
    TRY(dev_alloc(ft, &ft->d_pv_xyz, B * (size_t)ft->pv_stride));
    TRY(dev_alloc(ft, &ft->d_pv_start, B * (NP + 1)));
    if (ft->nn_f32) {
        TRY(dev_alloc(ft, &ft->d_pv_f32, B * (size_t)V));
        TRY(dev_alloc(ft, &ft->d_pv_rmax, B));
    }
    TRY(dev_alloc(ft, &ft->d_nn, NT));
    TRY(dev_alloc(ft, &ft->d_cnt, B * V));
    TRY(dev_alloc(ft, &ft->d_sum, B * 3 * V));
    TRY(dev_alloc(ft, &ft->d_range, B));
    TRY(dev_alloc(ft, &ft->d_Hcur, B * P * P));
    TRY(dev_alloc(ft, &ft->d_stats, B));
    for (int g : gnj) ft->max_nj = std::max(ft->max_nj, g);
    if (const char* e = std::getenv("AVB_FUSED")) ft->fused64 = std::atoi(e) != 0;
    if (const char* e = std::getenv("AVB_CHUNK")) ft->chunk_verts = ft->chunk_verts_tc = std::max(128, std::min(256, std::atoi(e) / 64 * 64));
    while (ft->chunk_verts > 128 && lm_gram_smem_bytes(ft->max_nj, m->K, ft->chunk_verts, false) > 112 * 1024) ft->chunk_verts -= 64;
    {
        const int cmin = std::min(ft->chunk_verts, ft->chunk_verts_tc);
        ft->maxc = (V + cmin - 1) / cmin + dp.numGroups + 1;
    }
    ft->tabD = lm_tab_doubles(m->J, m->K);
    ft->pstride = lm_part_stride(ft->max_nj, m->J, m->K);
    TRY(dev_alloc(ft, &ft->d_xt, B * nx));
    TRY(dev_alloc(ft, &ft->d_tab, B * (size_t)ft->tabD));
    TRY(dev_alloc(ft, &ft->d_part, B * (size_t)ft->maxc * (size_t)ft->pstride));
    ft->rec_rs = lm_rec_slots(V);
    ft->maxrb = std::max((ft->rec_rs + 255) / 256, ft->maxc);   // cost partials: one per record block (fp64 path) or per chunk (tensor path)
    ft->rec_stride = lm_rec_floats(ft->max_nj, m->K);
    TRY(dev_alloc(ft, &ft->d_cpart, B * (size_t)ft->maxrb));
    TRY(dev_alloc(ft, &ft->d_rec, B * (size_t)ft->rec_rs * (size_t)ft->rec_stride));
    TRY(dev_alloc(ft, &ft->d_gstart, B * (size_t)(kMaxGroups + 1)));
    TRY(dev_alloc(ft, &ft->d_gcur, B * P));
    TRY(dev_alloc(ft, &ft->d_mlist, B * (size_t)ft->rec_rs));
    TRY(dev_alloc(ft, &ft->d_chunks, B * (size_t)ft->maxc));
    TRY(dev_alloc(ft, &ft->d_gruns, B * (size_t)kMaxGroups));
    TRY(dev_alloc(ft, &ft->d_state, B));
    {   // work queue: ring of 4x the tasks that can be outstanding at once
        size_t need = 4 * B * (size_t)std::max(ft->maxrb, ft->maxc);
        ft->qcap = 1024;
        while (ft->qcap < need) ft->qcap <<= 1;
        TRY(dev_alloc(ft, &ft->d_qslots, (size_t)ft->qcap));
        TRY(dev_alloc(ft, &ft->d_qctrl, 8));
        TRY(dev_alloc(ft, &ft->d_rows_left, B));
        TRY(dev_alloc(ft, &ft->d_gram_left, B));
        ft->prior_stride = ((std::max(m->gmmD, 1) + 8) & ~7) + 8;
        TRY(dev_alloc(ft, &ft->d_prior_out, B * (size_t)ft->prior_stride));
        if (const char* e = std::getenv("AVB_PRIOR_TASK")) ft->prior_task_max = std::atoi(e);   // largest batch that runs the prior as its own task (0: never)
        TRY(dev_alloc(ft, &ft->d_qprof, 16));
        CUDA_TRY_FT(cudaMemset(ft->d_qslots, 0xFF, (size_t)ft->qcap * 8));
        CUDA_TRY_FT(cudaMemset(ft->d_qctrl, 0, 32));
        CUDA_TRY_FT(cudaMemset(ft->d_qprof, 0, 128));
        if (const char* e = std::getenv("AVB_FLOW")) ft->use_flow = std::atoi(e) != 0;
    }
    ft->max_chunks = (int)std::max<size_t>(NT / 512 + 2 * B + 8, (size_t)8 * ft->num_sms + 2 * B + 8);
    ft->max_qblocks = (int64_t)(NT / kQBlock + B + 8);
    TRY(dev_alloc(ft, &ft->d_qpart, (size_t)ft->max_qblocks));
    TRY(dev_alloc(ft, &ft->d_chunk_frame, (size_t)ft->max_chunks));
    TRY(dev_alloc(ft, &ft->d_chunk_count, (size_t)ft->max_chunks));
    TRY(dev_alloc(ft, &ft->d_chunk_qblock, (size_t)ft->max_chunks));
    TRY(dev_alloc(ft, &ft->d_chunk_begin, (size_t)ft->max_chunks));
    TRY(dev_alloc(ft, &ft->d_frame_qblock, B + 1));
    TRY(pin_alloc(ft, &ft->h_x, B * nx));
    TRY(pin_alloc(ft, &ft->h_stats, B));
    TRY(pin_alloc(ft, &ft->h_qctrl, 4));
    TRY(pin_alloc(ft, &ft->h_chunk_frame, (size_t)ft->max_chunks));
    TRY(pin_alloc(ft, &ft->h_chunk_count, (size_t)ft->max_chunks));
    TRY(pin_alloc(ft, &ft->h_chunk_qblock, (size_t)ft->max_chunks));
    TRY(pin_alloc(ft, &ft->h_chunk_begin, (size_t)ft->max_chunks));
    TRY(pin_alloc(ft, &ft->h_frame_qblock, B + 1));
    CUDA_TRY_FT(cudaMemsetAsync(ft->d_stats, 0, B * sizeof(FrameStats), ft->stream));
    CUDA_TRY_FT(cudaStreamSynchronize(ft->stream));
#undef TRY
#undef CUDA_TRY_FT
    *out = ft;
    return AVB_OK;
}

/* -------- batch upload: data + the NN chunk schedule -------- */
namespace {
// NN chunk schedule of a batch whose data are (or are about to be) resident: every frame is cut into equal chunks
// (multiples of 512 points) so that the grid has about 4 CTAs per SM; |d|^2 partials are always per 256 points,
// independent of the cut.  Enqueues the schedule tables on the fitter's stream.
int schedule_batch(avb_fitter* ft, int32_t batch, const int64_t* offsets) {
    const int64_t total = offsets[batch] - offsets[0];
    ft->offsets.assign(offsets, offsets + batch + 1);
    for (int f = 0; f < batch; ++f)
        if (offsets[f + 1] < offsets[f] || offsets[f + 1] - offsets[f] > (int64_t)1 << 21)
            return fail(AVB_ERR_INVALID, "offsets must be non-decreasing and a frame may hold at most 2^21 points");
    const int64_t o0 = offsets[0];
    int64_t target = std::max<int64_t>(512, (total / (4 * (int64_t)ft->num_sms) + 511) / 512 * 512);
    int nc = 0;
    int qb = 0;
    for (int f = 0; f < batch; ++f) {
        const int64_t n = offsets[f + 1] - offsets[f];
        ft->h_frame_qblock[f] = qb;
        if (n > 0) {
            const int64_t k = (n + target - 1) / target;
            const int64_t per = ((n + k - 1) / k + 511) / 512 * 512;
            for (int64_t b = 0; b < n; b += per) {
                if (nc >= ft->max_chunks) return fail(AVB_ERR_CAPACITY, "too many NN chunks");
                ft->h_chunk_frame[nc] = f;
                ft->h_chunk_begin[nc] = offsets[f] - o0 + b;
                ft->h_chunk_count[nc] = (int)std::min<int64_t>(per, n - b);
                ft->h_chunk_qblock[nc] = qb + (int)(b / kQBlock);
                ++nc;
            }
            qb += (int)((n + kQBlock - 1) / kQBlock);
        }
    }
    ft->h_frame_qblock[batch] = qb;
    if (qb > ft->max_qblocks) return fail(AVB_ERR_CAPACITY, "too many |d|^2 blocks");
    ft->batch = batch;
    ft->num_chunks = nc;
    ft->total_points = total;
    cudaStream_t st = ft->stream;
    if (nc > 0) {
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_frame, ft->h_chunk_frame, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_begin, ft->h_chunk_begin, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_count, ft->h_chunk_count, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_qblock, ft->h_chunk_qblock, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(ft->d_frame_qblock, ft->h_frame_qblock, (size_t)(batch + 1) * 4, cudaMemcpyHostToDevice, st));
    return AVB_OK;
}
}  // namespace

int avb_upload_batch(avb_fitter* ft, int32_t batch, const double* clouds, const int32_t* labels, const int64_t* offsets) {
    if (!ft || !offsets || batch <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    const int64_t total = offsets[batch] - offsets[0];
    if (total < 0 || total > ft->max_points) return fail(AVB_ERR_CAPACITY, "point count exceeds fitter capacity");
    if (total > 0 && (!clouds || !labels)) return fail(AVB_ERR_INVALID, "null cloud or labels");
    CUDA_TRY(cudaSetDevice(ft->device));
    int rc = schedule_batch(ft, batch, offsets);
    if (rc != AVB_OK) return rc;
    if (total > 0) {
        const int64_t o0 = offsets[0];
        CUDA_TRY(cudaMemcpyAsync(ft->d_data, clouds + 3 * o0, (size_t)total * 24, cudaMemcpyHostToDevice, ft->stream));
        ft->data_is_f32 = false;
        CUDA_TRY(cudaMemcpyAsync(ft->d_labels, labels + o0, (size_t)total * 4, cudaMemcpyHostToDevice, ft->stream));
    }
    return AVB_OK;
}

int avb_upload_batch_f32(avb_fitter* ft, int32_t batch, const float* clouds, const int32_t* labels, const int64_t* offsets) {
    if (!ft || !offsets || batch <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    const int64_t total = offsets[batch] - offsets[0];
    if (total < 0 || total > ft->max_points) return fail(AVB_ERR_CAPACITY, "point count exceeds fitter capacity");
    if (total > 0 && (!clouds || !labels)) return fail(AVB_ERR_INVALID, "null cloud or labels");
    CUDA_TRY(cudaSetDevice(ft->device));
    if (!ft->d_data_f32) {   // float staging of the cloud, allocated on first use (16-byte aligned by cudaMalloc)
        int rc0 = dev_alloc(ft, &ft->d_data_f32, 3 * (size_t)ft->max_points + 4);
        if (rc0 != AVB_OK) return rc0;
    }
    int rc = schedule_batch(ft, batch, offsets);
    if (rc != AVB_OK) return rc;
    if (total > 0) {
        const int64_t o0 = offsets[0];
        CUDA_TRY(cudaMemcpyAsync(ft->d_data_f32, clouds + 3 * o0, (size_t)total * 12, cudaMemcpyHostToDevice, ft->stream));
        CUDA_TRY(cudaMemcpyAsync(ft->d_labels, labels + o0, (size_t)total * 4, cudaMemcpyHostToDevice, ft->stream));
        ft->data_is_f32 = true;   // nn_kernel reads the floats and widens on load (exact): no separate pass over the cloud
    }
    return AVB_OK;
}

namespace {
// image staging of avb_upload_depth_batch / avb_rtree_predict_batch, grown on demand
int ensure_image_staging(avb_fitter* ft, size_t npx, size_t strip_slots) {
    if (npx <= ft->img_cap && strip_slots <= ft->strip_cap) return AVB_OK;
    npx = std::max(npx, ft->img_cap);
    strip_slots = std::max(strip_slots, ft->strip_cap);
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    cudaFree(ft->d_depth); cudaFree(ft->d_parts); cudaFree(ft->d_roi); cudaFree(ft->d_strip_count);
    cudaFree(ft->d_strip_offset); cudaFree(ft->d_bad_label);
    ft->d_depth = nullptr; ft->d_parts = nullptr; ft->d_roi = nullptr; ft->d_strip_count = nullptr;
    ft->d_strip_offset = nullptr; ft->d_bad_label = nullptr;
    ft->img_cap = ft->strip_cap = 0;
    if (cudaMalloc(&ft->d_depth, npx * 4) != cudaSuccess || cudaMalloc(&ft->d_parts, npx) != cudaSuccess ||
        cudaMalloc(&ft->d_roi, (size_t)ft->max_batch * 16) != cudaSuccess ||
        cudaMalloc(&ft->d_strip_count, std::max<size_t>(strip_slots, 1) * 4) != cudaSuccess ||
        cudaMalloc(&ft->d_strip_offset, std::max<size_t>(strip_slots, 1) * 8) != cudaSuccess ||
        cudaMalloc(&ft->d_bad_label, (size_t)ft->max_batch * 4) != cudaSuccess) {
        cudaGetLastError();
        return fail(AVB_ERR_CUDA, "cudaMalloc of the image staging buffers failed");
    }
    ft->img_cap = npx;
    ft->strip_cap = strip_slots;
    return AVB_OK;
}

// A bounding box handed to the RTree kernels must lie inside the image: the probes are looked up "inside the box"
// (RTree.cpp:52-67), so a box that leaves the image would read a neighbouring row or frame.  Empty boxes are fine.
int validate_rtree_roi(const int32_t* roi, int batch, int width, int height) {
    if (!roi) return AVB_OK;
    for (int f = 0; f < batch; ++f) {
        const int x0 = roi[4 * f], y0 = roi[4 * f + 1], x1 = roi[4 * f + 2], y1 = roi[4 * f + 3];
        if (x1 < x0 || y1 < y0) continue;   // empty box: nothing is predicted
        if (x0 < 0 || y0 < 0 || x1 >= width || y1 >= height)
            return fail(AVB_ERR_INVALID, "RTree bounding box of frame " + std::to_string(f) +
                                             " leaves the image (need 0 <= x0 <= x1 < width, 0 <= y0 <= y1 < height; inclusive corners)");
    }
    return AVB_OK;
}

// RTree::predictBest + upscaleGrid on the depth images resident in d_depth -> d_parts (roi already in d_roi if given)
int enqueue_rtree(avb_fitter* ft, int batch, int width, int height, const int32_t* roi_host, bool roi_on_device, int interval,
                  bool fill) {
    if (!ft->d_rt_nodes) return fail(AVB_ERR_INVALID, "no decision tree: call avb_fitter_set_rtree first");
    if (interval <= 0) return fail(AVB_ERR_INVALID, "RTree interval must be positive");
    cudaStream_t st = ft->stream;
    int max_w = width, max_h = height;
    if (roi_host) {
        max_w = max_h = 0;
        for (int f = 0; f < batch; ++f) {
            max_w = std::max(max_w, roi_host[4 * f + 2] - roi_host[4 * f] + 1);
            max_h = std::max(max_h, roi_host[4 * f + 3] - roi_host[4 * f + 1] + 1);
        }
        max_w = std::min(std::max(max_w, 0), width + interval);
        max_h = std::min(std::max(max_h, 0), height + interval);
    }
    RTreeArgs a{};
    a.nodes = ft->d_rt_nodes;
    a.leaf_best = ft->d_rt_leaf;
    a.depth = ft->d_depth;
    a.parts = ft->d_parts;
    a.roi = roi_on_device ? ft->d_roi : nullptr;
    a.width = width; a.height = height; a.interval = interval;
    if (!ft->rev[0])
        for (auto& e : ft->rev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaMemsetAsync(ft->d_parts, 0xFF, (size_t)batch * width * height, st));   // result.setTo(255)
    CUDA_TRY(cudaEventRecord(ft->rev[0], st));
    const long long cells = (long long)((max_w + interval - 1) / interval) * (max_h / interval + 1);
    CUDA_TRY(launch_rtree_predict(a, batch, (int)std::min<long long>(cells, 1 << 30), st));
    if (fill && interval > 1) {
        const long long px = (long long)(max_w + interval) * (max_h + 1);
        CUDA_TRY(launch_rtree_upscale(a, batch, (int)std::min<long long>(px, 1 << 30), st));
    }
    CUDA_TRY(cudaEventRecord(ft->rev[1], st));
    return AVB_OK;
}
}  // namespace

namespace {
int enqueue_postprocess(avb_fitter* ft, int batch, int width, int height, const int32_t* roi_host, int interval, int num_parts,
                        int part_map_type, double dist_w);
int ensure_com_state(avb_fitter* ft, int num_parts, bool reset);
}  // namespace

/* -------- cloud construction on the device (demo.cpp:215-250, Calibration.cpp:83-95) -------- */
int avb_upload_depth_batch(avb_fitter* ft, int32_t batch, const float* depth, const uint8_t* parts, const int32_t* roi,
                           const avb_image_desc* img, int64_t* offsets_out) {
    if (!ft || !depth || !img || batch <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (!parts && !ft->d_rt_nodes) return fail(AVB_ERR_INVALID, "parts == NULL needs a decision tree (avb_fitter_set_rtree)");
    if (!parts && img->rtree_interval <= 0) return fail(AVB_ERR_INVALID, "rtree_interval must be positive when parts == NULL");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    if (img->width <= 0 || img->height <= 0 || img->interval <= 0) return fail(AVB_ERR_INVALID, "bad image description");
    if (img->num_parts <= 0 || img->num_parts > 255) return fail(AVB_ERR_INVALID, "num_parts must be in 1..255");
    if (!(img->fx != 0.f) || !(img->fy != 0.f)) return fail(AVB_ERR_INVALID, "focal lengths must be non-zero");
    if (!parts) {
        const int rcv = validate_rtree_roi(roi, batch, img->width, img->height);
        if (rcv != AVB_OK) return rcv;
    }
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const size_t px = (size_t)img->width * img->height, npx = px * batch;
    const int rows_per_strip = cloud_strip_rows();
    const int strips = ((img->height + img->interval - 1) / img->interval + rows_per_strip - 1) / rows_per_strip;
    {
        int rc0 = ensure_image_staging(ft, npx, (size_t)ft->max_batch * strips);
        if (rc0 != AVB_OK) return rc0;
    }
    CUDA_TRY(cudaMemcpyAsync(ft->d_depth, depth, npx * 4, cudaMemcpyHostToDevice, st));
    if (roi) CUDA_TRY(cudaMemcpyAsync(ft->d_roi, roi, (size_t)batch * 16, cudaMemcpyHostToDevice, st));
    if (parts) {
        CUDA_TRY(cudaMemcpyAsync(ft->d_parts, parts, npx, cudaMemcpyHostToDevice, st));
    } else {   // demo.cpp:198-200: labels from the decision tree, on the device
        int rcr = enqueue_rtree(ft, batch, img->width, img->height, roi, roi != nullptr, img->rtree_interval, true);
        if (rcr != AVB_OK) return rcr;
        if (img->rtree_postprocess) {   // demo.cpp:201: rtree.postProcess(result, comPre, 2, threads, topLeft, botRight)
            rcr = ensure_com_state(ft, img->num_parts, false);
            if (rcr == AVB_OK)
                rcr = enqueue_postprocess(ft, batch, img->width, img->height, roi, img->rtree_interval, img->num_parts,
                                          img->part_map_type, img->dist_to_pre_weight > 0.0 ? img->dist_to_pre_weight : 0.001);
            if (rcr != AVB_OK) return rcr;
        }
    }
    CUDA_TRY(cudaMemsetAsync(ft->d_bad_label, 0, (size_t)batch * 4, st));
    CloudArgs a{};
    a.depth = ft->d_depth;
    a.parts = ft->d_parts;
    a.roi = roi ? ft->d_roi : nullptr;
    a.width = img->width; a.height = img->height; a.interval = img->interval; a.num_parts = img->num_parts;
    a.fx = img->fx; a.cx = img->cx; a.fy = img->fy; a.cy = img->cy;
    a.max_strips = strips;
    a.strip_count = ft->d_strip_count;
    a.strip_offset = ft->d_strip_offset;
    a.bad_label = ft->d_bad_label;
    a.cloud = ft->d_data;
    ft->data_is_f32 = false;   // the cloud kernels write doubles
    a.labels = ft->d_labels;
    if (!ft->cev[0])
        for (auto& e : ft->cev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ft->cev[0], st));
    CUDA_TRY(launch_cloud_count(a, strips, batch, st));
    CUDA_TRY(cudaEventRecord(ft->cev[1], st));
    // strip counts -> host: frame offsets, strip offsets and the NN schedule (one small round trip per batch)
    ft->h_strip_count.resize((size_t)batch * strips + batch);
    ft->h_strip_offset.resize((size_t)batch * strips);
    CUDA_TRY(cudaMemcpyAsync(ft->h_strip_count.data(), ft->d_strip_count, (size_t)batch * strips * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ft->h_strip_count.data() + (size_t)batch * strips, ft->d_bad_label, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int f = 0; f < batch; ++f)
        if (ft->h_strip_count[(size_t)batch * strips + f])
            return fail(AVB_ERR_INVALID, "body-part label >= num_parts in frame " + std::to_string(f) +
                                             " (the reference exits here, demo.cpp:232-239)");
    std::vector<int64_t> off((size_t)batch + 1, 0);
    for (int f = 0; f < batch; ++f) {
        int64_t run = off[f];
        for (int s2 = 0; s2 < strips; ++s2) {
            ft->h_strip_offset[(size_t)f * strips + s2] = run;
            run += ft->h_strip_count[(size_t)f * strips + s2];
        }
        off[f + 1] = run;
    }
    if (off[batch] > ft->max_points) return fail(AVB_ERR_CAPACITY, "point count exceeds fitter capacity");
    int rc = schedule_batch(ft, batch, off.data());
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(ft->d_strip_offset, ft->h_strip_offset.data(), (size_t)batch * strips * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(ft->cev[2], st));
    CUDA_TRY(launch_cloud_compact(a, strips, batch, st));
    CUDA_TRY(cudaEventRecord(ft->cev[3], st));
    if (offsets_out) std::copy(off.begin(), off.end(), offsets_out);
    return AVB_OK;
}

int avb_fitter_set_rtree(avb_fitter* ft, const avb_rtree_desc* t) {
    if (!ft || !t) return fail(AVB_ERR_INVALID, "null argument");
    if (t->num_nodes <= 0 || t->num_leaves <= 0 || !t->u || !t->v || !t->thresh || !t->lnode || !t->rnode || !t->leafid || !t->leaf_best)
        return fail(AVB_ERR_INVALID, "incomplete decision tree");
    std::vector<RTreeNode> nodes((size_t)t->num_nodes);
    for (int i = 0; i < t->num_nodes; ++i) {
        RTreeNode& n = nodes[i];
        n.ux = t->u[2 * i]; n.uy = t->u[2 * i + 1]; n.vx = t->v[2 * i]; n.vy = t->v[2 * i + 1];
        n.thresh = t->thresh[i];
        n.lnode = t->lnode[i]; n.rnode = t->rnode[i]; n.leafid = t->leafid[i];
        if (n.leafid < -1 || n.leafid >= t->num_leaves) return fail(AVB_ERR_INVALID, "leaf id out of range");
        if (n.leafid == -1 && (n.lnode <= i || n.lnode >= t->num_nodes || n.rnode <= i || n.rnode >= t->num_nodes))
            return fail(AVB_ERR_INVALID, "child index out of range (children must follow their parent)");
    }
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    cudaFree(ft->d_rt_nodes); cudaFree(ft->d_rt_leaf);   // only the previous tree: the renderer buffers are not this call's
    ft->d_rt_nodes = nullptr; ft->d_rt_leaf = nullptr;
    ft->rt_nodes = ft->rt_leaves = 0;
    CUDA_TRY(cudaMalloc(&ft->d_rt_nodes, nodes.size() * sizeof(RTreeNode)));
    CUDA_TRY(cudaMalloc(&ft->d_rt_leaf, (size_t)t->num_leaves));
    CUDA_TRY(cudaMemcpy(ft->d_rt_nodes, nodes.data(), nodes.size() * sizeof(RTreeNode), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ft->d_rt_leaf, t->leaf_best, (size_t)t->num_leaves, cudaMemcpyHostToDevice));
    ft->rt_nodes = t->num_nodes; ft->rt_leaves = t->num_leaves; ft->rt_parts = t->num_parts;
    return AVB_OK;
}

namespace {
// RTree::postProcess on the label images resident in d_parts; com_pre on the device (ft->d_pp_com), roi in d_roi if given
int enqueue_postprocess(avb_fitter* ft, int batch, int width, int height, const int32_t* roi_host, int interval, int num_parts,
                        int part_map_type, double dist_w) {
    if (interval <= 0) return fail(AVB_ERR_INVALID, "interval must be positive");
    if (num_parts <= 0 || num_parts > 64) return fail(AVB_ERR_INVALID, "postProcess handles 1..64 parts (labels < 128)");
    if (width > 65535 || height > 32767) return fail(AVB_ERR_INVALID, "postProcess packs pixel ids into 16 + 15 bits");
    long long grid = 0;
    for (int f = 0; f < batch; ++f) {
        const int x0 = roi_host ? roi_host[4 * f] : 0, y0 = roi_host ? roi_host[4 * f + 1] : 0;
        const int x1 = roi_host ? roi_host[4 * f + 2] : width - 1, y1 = roi_host ? roi_host[4 * f + 3] : height - 1;
        if (x1 < x0 || y1 < y0) continue;
        if (x0 < 0 || y0 < 0 || x1 >= width || y1 >= height)
            return fail(AVB_ERR_INVALID, "postProcess bounding box of frame " + std::to_string(f) + " leaves the image");
        grid = std::max(grid, (long long)((x1 - x0) / interval + 1) * ((y1 - y0) / interval + 1));
    }
    const size_t cap = (size_t)(2 * grid + 64);
    if (cap > ft->pp_cap) {
        CUDA_TRY(cudaStreamSynchronize(ft->stream));
        cudaFree(ft->d_pp_arena); cudaFree(ft->d_pp_stack);
        ft->d_pp_arena = ft->d_pp_stack = nullptr;
        ft->pp_cap = 0;
        CUDA_TRY(cudaMalloc(&ft->d_pp_arena, (size_t)ft->max_batch * cap * 4));
        CUDA_TRY(cudaMalloc(&ft->d_pp_stack, (size_t)ft->max_batch * cap * 4));
        ft->pp_cap = cap;
    }
    if (!ft->d_pp_ovf) CUDA_TRY(cudaMalloc(&ft->d_pp_ovf, (size_t)ft->max_batch * 4));
    CUDA_TRY(cudaMemsetAsync(ft->d_pp_ovf, 0, (size_t)batch * 4, ft->stream));
    RTreePostArgs a{};
    a.parts = ft->d_parts;
    a.roi = roi_host ? ft->d_roi : nullptr;
    a.width = width; a.height = height; a.interval = interval; a.num_parts = num_parts; a.part_map_type = part_map_type;
    a.dist_w = dist_w;
    a.com_pre = ft->d_pp_com;
    a.arena = ft->d_pp_arena;
    a.stack = ft->d_pp_stack;
    a.cap = (long long)ft->pp_cap;
    a.overflow = ft->d_pp_ovf;
    if (!ft->rev[0])
        for (auto& e : ft->rev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ft->rev[0], ft->stream));   // avb_last_rtree_ms after a stand-alone postProcess call: this kernel
    CUDA_TRY(launch_rtree_postprocess(a, batch, ft->stream));
    CUDA_TRY(cudaEventRecord(ft->rev[1], ft->stream));
    return AVB_OK;
}
int ensure_com_state(avb_fitter* ft, int num_parts, bool reset) {
    if (!ft->d_pp_com || ft->pp_parts != num_parts) {
        CUDA_TRY(cudaStreamSynchronize(ft->stream));
        cudaFree(ft->d_pp_com);
        ft->d_pp_com = nullptr;
        CUDA_TRY(cudaMalloc(&ft->d_pp_com, (size_t)ft->max_batch * 2 * num_parts * 8));
        ft->pp_parts = num_parts;
        reset = true;
    }
    if (reset) {   // the reference's freshly resized com_pre: x = -1 (not seen), y = 0 (RTree.cpp:3432-3436)
        std::vector<double> init((size_t)ft->max_batch * 2 * num_parts, 0.0);
        for (size_t i = 0; i < init.size(); i += 2) init[i] = -1.0;
        CUDA_TRY(cudaMemcpyAsync(ft->d_pp_com, init.data(), init.size() * 8, cudaMemcpyHostToDevice, ft->stream));
        CUDA_TRY(cudaStreamSynchronize(ft->stream));
    }
    return AVB_OK;
}
}  // namespace

/* RTree::postProcess (RTree.cpp:3422-3450) on a batch of label images */
int avb_rtree_postprocess_batch(avb_fitter* ft, int32_t batch, uint8_t* parts, int32_t width, int32_t height, const int32_t* roi,
                                int32_t interval, int32_t num_parts, int32_t part_map_type, double* com_pre, double dist_to_pre_weight) {
    if (!ft || !parts || !com_pre || batch <= 0 || width <= 0 || height <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    if (num_parts <= 0 || num_parts > 64) return fail(AVB_ERR_INVALID, "postProcess handles 1..64 parts (labels < 128)");
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t npx = (size_t)width * height * batch;
    int rc = ensure_image_staging(ft, npx, 0);
    if (rc != AVB_OK) return rc;
    rc = ensure_com_state(ft, num_parts, false);
    if (rc != AVB_OK) return rc;
    cudaStream_t st = ft->stream;
    CUDA_TRY(cudaMemcpyAsync(ft->d_parts, parts, npx, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ft->d_pp_com, com_pre, (size_t)batch * 2 * num_parts * 8, cudaMemcpyHostToDevice, st));
    if (roi) CUDA_TRY(cudaMemcpyAsync(ft->d_roi, roi, (size_t)batch * 16, cudaMemcpyHostToDevice, st));
    rc = enqueue_postprocess(ft, batch, width, height, roi, interval, num_parts, part_map_type, dist_to_pre_weight);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(parts, ft->d_parts, npx, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(com_pre, ft->d_pp_com, (size_t)batch * 2 * num_parts * 8, cudaMemcpyDeviceToHost, st));
    std::vector<int> ovf((size_t)batch);
    CUDA_TRY(cudaMemcpyAsync(ovf.data(), ft->d_pp_ovf, (size_t)batch * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int f = 0; f < batch; ++f)
        if (ovf[f]) return fail(AVB_ERR_CAPACITY, "postProcess scratch overflow in frame " + std::to_string(f) + " (internal error)");
    return AVB_OK;
}

int avb_rtree_reset_tracking(avb_fitter* ft) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    CUDA_TRY(cudaSetDevice(ft->device));
    if (ft->pp_parts > 0) return ensure_com_state(ft, ft->pp_parts, true);
    return AVB_OK;
}

int avb_rtree_predict_batch(avb_fitter* ft, int32_t batch, const float* depth, int32_t width, int32_t height, const int32_t* roi,
                            int32_t interval, int32_t fill_in_gaps, uint8_t* parts_out) {
    if (!ft || !depth || !parts_out || batch <= 0 || width <= 0 || height <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    if (!ft->d_rt_nodes) return fail(AVB_ERR_INVALID, "no decision tree: call avb_fitter_set_rtree first");
    if (interval <= 0) return fail(AVB_ERR_INVALID, "RTree interval must be positive");
    {
        const int rcv = validate_rtree_roi(roi, batch, width, height);
        if (rcv != AVB_OK) return rcv;
    }
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t npx = (size_t)width * height * batch;
    int rc = ensure_image_staging(ft, npx, 0);
    if (rc != AVB_OK) return rc;
    cudaStream_t st = ft->stream;
    CUDA_TRY(cudaMemcpyAsync(ft->d_depth, depth, npx * 4, cudaMemcpyHostToDevice, st));
    if (roi) CUDA_TRY(cudaMemcpyAsync(ft->d_roi, roi, (size_t)batch * 16, cudaMemcpyHostToDevice, st));
    rc = enqueue_rtree(ft, batch, width, height, roi, roi != nullptr, interval, fill_in_gaps != 0);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(parts_out, ft->d_parts, npx, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AVB_OK;
}

int avb_last_rtree_ms(avb_fitter* ft, float* ms) {
    if (!ft || !ms) return fail(AVB_ERR_INVALID, "null argument");
    if (!ft->rev[0]) return fail(AVB_ERR_INVALID, "no RTree prediction yet");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventSynchronize(ft->rev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, ft->rev[0], ft->rev[1]));
    return AVB_OK;
}

int avb_last_cloud_ms(avb_fitter* ft, float* ms2) {
    if (!ft || !ms2) return fail(AVB_ERR_INVALID, "null argument");
    if (!ft->cev[0]) return fail(AVB_ERR_INVALID, "no avb_upload_depth_batch yet");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventSynchronize(ft->cev[3]));
    CUDA_TRY(cudaEventElapsedTime(&ms2[0], ft->cev[0], ft->cev[1]));
    CUDA_TRY(cudaEventElapsedTime(&ms2[1], ft->cev[2], ft->cev[3]));
    return AVB_OK;
}

int avb_download_batch(avb_fitter* ft, double* clouds, int32_t* labels, int64_t* offsets) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    if (ft->batch <= 0) return fail(AVB_ERR_INVALID, "no batch uploaded");
    CUDA_TRY(cudaSetDevice(ft->device));
    if (clouds && ft->total_points > 0) {
        if (ft->data_is_f32)   // the resident cloud is the float upload: widen it for the caller
            CUDA_TRY(launch_widen_points(ft->d_data_f32, ft->d_data, 3 * (long long)ft->total_points, ft->num_sms, ft->stream));
        CUDA_TRY(cudaMemcpyAsync(clouds, ft->d_data, (size_t)ft->total_points * 24, cudaMemcpyDeviceToHost, ft->stream));
    }
    if (labels && ft->total_points > 0)
        CUDA_TRY(cudaMemcpyAsync(labels, ft->d_labels, (size_t)ft->total_points * 4, cudaMemcpyDeviceToHost, ft->stream));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    if (offsets) std::copy(ft->offsets.begin(), ft->offsets.end(), offsets);
    return AVB_OK;
}

namespace {

int check_options(const avb_fitter* ft, const avb_options* o) {
    if (!o) return fail(AVB_ERR_INVALID, "null options");
    if (o->icp_iters < 0 || o->max_iters_per_icp < 0) return fail(AVB_ERR_INVALID, "negative iteration count");
    if (o->solver != AVB_SOLVER_GN_LM) return fail(AVB_ERR_INVALID, "unknown solver");
    if (o->jtj_precision != AVB_JTJ_FP64 && o->jtj_precision != AVB_JTJ_FP32 && o->jtj_precision != AVB_JTJ_BF16_TENSOR)
        return fail(AVB_ERR_INVALID, "jtj_precision must be AVB_JTJ_FP64, AVB_JTJ_FP32 or AVB_JTJ_BF16_TENSOR");
    if (o->jtj_precision == AVB_JTJ_BF16_TENSOR && !lm_tensor_supported(ft->max_nj, ft->model->K))
        return fail(AVB_ERR_INVALID, "AVB_JTJ_BF16_TENSOR needs at most 80 record fields (3 joints + 1 + 3 shape keys) and 64 Jacobian "
                                     "columns per column group: use AVB_JTJ_FP64 for this model (or more groups, AVB_GROUPS)");
    if (o->beta_pose > 0.0 && ft->model->gmmC <= 0)
        return fail(AVB_ERR_PRIOR, "betaPose > 0 but the model has no pose prior");
    return AVB_OK;
}

// kernel classes for avb_last_kernel_ms
enum { KC_POSE = 0, KC_NN = 1, KC_PREP = 2, KC_ROWS = 3, KC_GRAM = 4, KC_SOLVE = 5, KC_FINAL = 6, KC_FLOW = 7, KC_COUNT = 8 };
struct ProfScope {   // records an event pair around one launch when profiling is on
    avb_fitter* ft;
    size_t idx;
    bool on;
    ProfScope(avb_fitter* f, int cls) : ft(f), idx(0), on(f->profile) {
        if (!on) return;
        idx = ft->pcls.size();
        ft->pcls.push_back(cls);
        while (ft->pev.size() < 2 * (idx + 1)) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ft->pev.push_back(e);
        }
        cudaEventRecord(ft->pev[2 * idx], ft->stream);
    }
    ~ProfScope() {
        if (on) cudaEventRecord(ft->pev[2 * idx + 1], ft->stream);
    }
};

PoseArgs pose_args(avb_fitter* ft, const double* dx, bool vis, const avb_options* o) {
    PoseArgs a{};
    a.x = dx;
    a.cloud = ft->d_cloud;
    a.joint_pos = nullptr;
    a.joint_trans = nullptr;
    a.do_visibility = vis ? 1 : 0;
    a.do_lbs = 1;
    // small batches: a forward-only launch spreads the vertices of a frame over several CTAs (every CTA rebuilds the joint tables)
    a.slices = ft->batch * 8 <= ft->num_sms ? 8 : (ft->batch * 4 <= ft->num_sms ? 4 : (ft->batch * 2 <= ft->num_sms ? 2 : 1));
    a.enable_occlusion = o ? o->enable_occlusion : 1;
    a.visible = ft->d_vis;
    a.pv_idx = ft->d_pv_idx;
    a.pv_xyz = ft->d_pv_xyz;
    a.pv_start = ft->d_pv_start;
    a.pv_stride = ft->pv_stride;
    a.pv_f32 = vis ? ft->d_pv_f32 : nullptr;
    a.pv_rmax = ft->d_pv_rmax;
    return a;
}

int enqueue_correspond(avb_fitter* ft, const double* dx, const avb_options* o, cudaEvent_t after_pose) {
    cudaStream_t st = ft->stream;
    const int B = ft->batch, V = ft->model->V;
    PoseArgs pa = pose_args(ft, dx, true, o);
    pa.zero_cnt = ft->d_cnt;   // zeroed by the kernel's visibility pass: no memset launches between pose and nn
    pa.zero_sum = ft->d_sum;
    pa.zero_range = ft->d_range;
    {
        ProfScope ps(ft, KC_POSE);
        if (pa.slices > 1) {   // small batch: pose the cloud with several CTAs per frame, then visibility + compaction per frame
            PoseArgs fwd = pa;
            fwd.do_visibility = 0;
            fwd.pv_f32 = nullptr;
            fwd.zero_cnt = nullptr;
            CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, fwd, B, st));
            ++ft->launches;
            pa.do_lbs = 0;
        }
        CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, B, st));
    }
    ++ft->launches;
    if (after_pose) CUDA_TRY(cudaEventRecord(after_pose, st));
    NNArgs na{};
    na.V = V;
    na.data = ft->d_data;
    na.data_f32 = ft->data_is_f32 ? ft->d_data_f32 : nullptr;
    na.labels = ft->d_labels;
    na.chunk_frame = ft->d_chunk_frame;
    na.chunk_begin = ft->d_chunk_begin;
    na.chunk_count = ft->d_chunk_count;
    na.chunk_qblock = ft->d_chunk_qblock;
    na.pv_idx = ft->d_pv_idx;
    na.pv_xyz = ft->d_pv_xyz;
    na.pv_start = ft->d_pv_start;
    na.pv_stride = ft->pv_stride;
    na.pv_f32 = ft->d_pv_f32;
    na.pv_rmax = ft->d_pv_rmax;
    na.x = dx;
    na.nx = ft->model->nx;
    na.nn_idx = ft->d_nn;
    na.cnt = ft->d_cnt;
    na.sum = ft->d_sum;
    na.qpart = ft->d_qpart;
    na.range_flag = ft->d_range;
    na.stage_cap = ft->nn_stage_cap;
    {
        ProfScope ps(ft, KC_NN);
        CUDA_TRY(launch_nn(ft->dp, na, ft->num_chunks, st));
    }
    if (ft->num_chunks > 0) ++ft->launches;
    return AVB_OK;
}

LmBuf lm_buf(avb_fitter* ft, double* dx, const avb_options* o) {
    LmBuf a{};
    a.x = dx;
    a.xt = ft->d_xt;
    a.tab = ft->d_tab;
    a.mlist = ft->d_mlist;
    a.chunks = ft->d_chunks;
    a.gruns = ft->d_gruns;
    a.part = ft->d_part;
    a.cpart = ft->d_cpart;
    a.rec = ft->d_rec;
    a.gstart = ft->d_gstart;
    a.maxrb = ft->maxrb;
    a.rec_stride = ft->rec_stride;
    a.rec_rs = ft->rec_rs;
    a.gcur = ft->d_gcur;
    a.Hcur = ft->d_Hcur;
    a.state = ft->d_state;
    a.maxc = ft->maxc;
    a.tabD = ft->tabD;
    a.chunk_verts = o->jtj_precision == AVB_JTJ_BF16_TENSOR ? ft->chunk_verts_tc : ft->chunk_verts;   // the chunk tables are rebuilt by every lm_prep_kernel
    a.pstride = ft->pstride;
    a.cnt = ft->d_cnt;
    a.sum = ft->d_sum;
    a.qpart = ft->d_qpart;
    a.frame_qblock = ft->d_frame_qblock;
    a.range_flag = ft->d_range;
    a.beta_pose = o->beta_pose;
    a.beta_shape = o->beta_shape;
    a.function_tolerance = o->function_tolerance;
    a.max_iters = o->max_iters_per_icp;
    a.stats = ft->d_stats;
    // one persistent data-flow kernel for the whole inner solve (the staged kernels, AVB_FLOW=0, exist for the fp64 path only)
    a.tensor = o->jtj_precision == AVB_JTJ_BF16_TENSOR ? 1 : 0;
    if (ft->use_flow || a.tensor) {
        a.fused = a.tensor ? 0 : ft->fused64;
        a.q.slots = ft->d_qslots;
        a.q.ctrl = ft->d_qctrl;
        a.q.rows_left = ft->d_rows_left;
        a.q.gram_left = ft->d_gram_left;
        a.q.cap_mask = ft->qcap - 1;
        // small batches leave most SMs idle: the pose prior of a trial point runs as its own task next to the record / Gram tasks
        a.prior_task = (ft->batch <= ft->prior_task_max && o->beta_pose > 0.0 && ft->model->gmmC > 0) ? 1 : 0;
        a.prior_out = ft->d_prior_out;
        a.prior_stride = ft->prior_stride;
    }
    a.q.prof = ft->profile ? ft->d_qprof : nullptr;   // per-phase CTA time (also with the staged kernels)
    return a;
}

// the inner solve of one ICP iteration: prep + (1 + max_iters) evaluations, all asynchronous
int enqueue_solve(avb_fitter* ft, const LmBuf& la, const avb_options* o, int rounds) {
    cudaStream_t st = ft->stream;
    {
        ProfScope ps(ft, KC_PREP);
        CUDA_TRY(launch_lm_prep(ft->dm, ft->dp, la, ft->batch, st));
    }
    ++ft->launches;
    if (la.q.prof) CUDA_TRY(cudaMemsetAsync(ft->d_qprof, 0, 128, st));
    if (la.q.slots) {
        // CTAs per SM: what the shared memory of the variant allows (227 KB per SM, 1 KB reserved per CTA), capped by the
        // register budget the kernel variants are compiled for; AVB_FLOW_OCC overrides
        const size_t fsm = lm_flow_smem_bytes(ft->dm, ft->max_nj, la.chunk_verts, la.tensor != 0);
        const int occ_cap = (int)std::min<size_t>(la.tensor ? 4 : 3, (227 * 1024) / (fsm + 1024));
        int occ = ft->flow_occ_default[la.tensor ? 1 : 0];
        if (const char* e = std::getenv("AVB_FLOW_OCC")) occ = std::atoi(e);
        occ = std::max(2, std::min(occ, occ_cap));
        int ctas = std::min(occ * ft->num_sms, std::max(1, ft->batch * std::max(16, ft->maxc)));
        if (const char* e = std::getenv("AVB_FLOW_CTAS")) ctas = std::max(1, std::min(ctas, std::atoi(e)));
        ProfScope ps(ft, KC_FLOW);
        CUDA_TRY(launch_lm_flow(ft->dm, ft->dp, la, ft->max_nj, ctas, occ, st));
        ++ft->launches;
        (void)rounds;
        return AVB_OK;
    }
    const bool tensor = o->jtj_precision == AVB_JTJ_BF16_TENSOR;   // AVB_JTJ_FP32 permits, but no longer uses, fp32 sums
    for (int r = 0; r < rounds; ++r) {
        for (int part = 0; part < 3; ++part) {
            ProfScope ps(ft, KC_ROWS + part);
            CUDA_TRY(launch_lm_eval_part(ft->dm, ft->dp, la, ft->batch, ft->max_nj, tensor, part, st));
        }
        ft->launches += 3;
    }
    return AVB_OK;
}

}  // namespace

int avb_fit_resident(avb_fitter* ft, const double* x_in, const avb_options* o) {
    if (!ft || !x_in) return fail(AVB_ERR_INVALID, "null argument");
    if (ft->batch <= 0) return fail(AVB_ERR_INVALID, "no batch uploaded");
    int rc = check_options(ft, o);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const int B = ft->batch;
    const size_t nx = ft->model->nx;
    // h_x is the one pinned staging buffer of the start points: the copy of the previous (still queued) call must have
    // read it before it is rewritten, or back-to-back fits from different warm starts would see each other's x
    if (ft->hx_busy) CUDA_TRY(cudaEventSynchronize(ft->hx_ev));
    std::memcpy(ft->h_x, x_in, (size_t)B * nx * 8);
    CUDA_TRY(cudaMemcpyAsync(ft->d_x, ft->h_x, (size_t)B * nx * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(ft->hx_ev, st));
    ft->hx_busy = true;
    ft->launches = 0;
    ft->last_icp = o->icp_iters;
    ft->pcls.clear();
    CUDA_TRY(cudaEventRecord(ft->ev[0], st));
    for (int icp = 0; icp < o->icp_iters; ++icp) {
        const bool timed = (icp == o->icp_iters - 1);
        if (timed) CUDA_TRY(cudaEventRecord(ft->ev[1], st));
        rc = enqueue_correspond(ft, ft->d_x, o, timed ? ft->ev[2] : nullptr);
        if (rc != AVB_OK) return rc;
        if (timed) CUDA_TRY(cudaEventRecord(ft->ev[3], st));
        LmBuf la = lm_buf(ft, ft->d_x, o);
        rc = enqueue_solve(ft, la, o, 1 + o->max_iters_per_icp);
        if (rc != AVB_OK) return rc;
        if (timed) CUDA_TRY(cudaEventRecord(ft->ev[4], st));
    }
    // trailing ava.update() (AvatarOptimizer.cpp:1497)
    PoseArgs pa = pose_args(ft, ft->d_x, false, o);
    {
        ProfScope ps(ft, KC_FINAL);
        CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, B, st));
    }
    ++ft->launches;
    CUDA_TRY(cudaEventRecord(ft->ev[5], st));
    return AVB_OK;
}

/* ---------------- multi-GPU entry points: one NCCL all-gather of the fitted parameters ---------------- */
int avb_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(AVB_ERR_INVALID, "null argument");
    if (!nccl().ok) return fail(AVB_ERR_CUDA, "libnccl.so.2 could not be loaded: the multi-GPU entry points need NCCL");
    NcclApi::UniqueId id;
    const int rc = nccl().GetUniqueId(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(id128, id.internal, 128);
    return AVB_OK;
}

int avb_fitter_comm_init(avb_fitter* ft, const uint8_t* id128, int32_t rank, int32_t nranks) {
    if (!ft || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return fail(AVB_ERR_INVALID, "bad communicator arguments");
    if (!nccl().ok) return fail(AVB_ERR_CUDA, "libnccl.so.2 could not be loaded: the multi-GPU entry points need NCCL");
    if (ft->nccl_comm) return fail(AVB_ERR_INVALID, "the fitter already has a communicator");
    CUDA_TRY(cudaSetDevice(ft->device));
    NcclApi::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    const int rc = nccl().CommInitRank(&ft->nccl_comm, nranks, id, rank);
    if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
    ft->comm_rank = rank;
    ft->comm_size = nranks;
    const size_t n = (size_t)nranks * ft->max_batch * ft->model->nx;
    int r2 = dev_alloc(ft, &ft->d_gather, n);
    if (r2 == AVB_OK) r2 = pin_alloc(ft, &ft->h_gather, n);
    if (r2 == AVB_OK) r2 = dev_alloc(ft, &ft->d_send, (size_t)ft->max_batch * ft->model->nx);
    if (r2 != AVB_OK) return r2;
    CUDA_TRY(cudaStreamCreateWithFlags(&ft->gather_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ft->ev_send, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ft->ev_gdone, cudaEventDisableTiming));
    return AVB_OK;
}

// The gather runs on its OWN stream behind a snapshot of d_x: the caller's host thread (and the fitter's stream) go on with
// the next batch while NCCL's kernel waits for an SM slot (the persistent lm_flow_kernel of a neighbouring fitter may hold
// every slot for milliseconds) and for the peers.  begin / end must alternate.
int avb_gather_params_begin(avb_fitter* ft) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    if (!ft->nccl_comm) return fail(AVB_ERR_INVALID, "no communicator: call avb_fitter_comm_init first");
    if (ft->gather_pending) return fail(AVB_ERR_INVALID, "a gather is already in flight: call avb_gather_params_end first");
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t per = (size_t)ft->max_batch * ft->model->nx;
    CUDA_TRY(cudaMemcpyAsync(ft->d_send, ft->d_x, per * 8, cudaMemcpyDeviceToDevice, ft->stream));
    CUDA_TRY(cudaEventRecord(ft->ev_send, ft->stream));
    CUDA_TRY(cudaStreamWaitEvent(ft->gather_stream, ft->ev_send, 0));
    const int rc = nccl().AllGather(ft->d_send, ft->d_gather, per, /*ncclFloat64*/ 8, ft->nccl_comm, ft->gather_stream);
    if (rc != 0) return nccl_fail("ncclAllGather", rc);
    CUDA_TRY(cudaMemcpyAsync(ft->h_gather, ft->d_gather, per * ft->comm_size * 8, cudaMemcpyDeviceToHost, ft->gather_stream));
    CUDA_TRY(cudaEventRecord(ft->ev_gdone, ft->gather_stream));
    // (the next snapshot cannot overwrite d_send early: begin / end alternate, and end waits for ev_gdone on the host.  No
    //  wait is put on the fitter's stream here -- it would chain the next upload and fit behind the collective.)
    ft->gather_pending = true;
    return AVB_OK;
}

int avb_gather_params_end(avb_fitter* ft, double* all_x) {
    if (!ft || !all_x) return fail(AVB_ERR_INVALID, "null argument");
    if (!ft->gather_pending) return fail(AVB_ERR_INVALID, "no gather in flight: call avb_gather_params_begin first");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventSynchronize(ft->ev_gdone));
    ft->gather_pending = false;
    std::memcpy(all_x, ft->h_gather, (size_t)ft->max_batch * ft->model->nx * ft->comm_size * 8);
    return AVB_OK;
}

int avb_gather_params(avb_fitter* ft, double* all_x) {
    if (!ft || !all_x) return fail(AVB_ERR_INVALID, "null argument");
    const int rc = avb_gather_params_begin(ft);
    return rc != AVB_OK ? rc : avb_gather_params_end(ft, all_x);
}

int avb_synchronize(avb_fitter* ft) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    return AVB_OK;
}

int avb_last_device_ms(avb_fitter* ft, float* total_ms, float* per4) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventSynchronize(ft->ev[5]));
    if (total_ms) CUDA_TRY(cudaEventElapsedTime(total_ms, ft->ev[0], ft->ev[5]));
    if (per4) {
        per4[0] = per4[1] = per4[2] = per4[3] = 0.f;
        if (ft->last_icp > 0) {
            CUDA_TRY(cudaEventElapsedTime(&per4[0], ft->ev[1], ft->ev[2]));
            CUDA_TRY(cudaEventElapsedTime(&per4[1], ft->ev[2], ft->ev[3]));
            CUDA_TRY(cudaEventElapsedTime(&per4[2], ft->ev[3], ft->ev[4]));
            CUDA_TRY(cudaEventElapsedTime(&per4[3], ft->ev[4], ft->ev[5]));
        }
    }
    return AVB_OK;
}

int avb_timer_start(avb_fitter* ft) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventRecord(ft->ev[6], ft->stream));
    return AVB_OK;
}
int avb_timer_stop(avb_fitter* ft, float* ms) {
    if (!ft || !ms) return fail(AVB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventRecord(ft->ev[7], ft->stream));
    CUDA_TRY(cudaEventSynchronize(ft->ev[7]));
    CUDA_TRY(cudaEventElapsedTime(ms, ft->ev[6], ft->ev[7]));
    return AVB_OK;
}

int avb_last_launch_count(avb_fitter* ft) { return ft ? ft->launches : 0; }

int avb_set_profiling(avb_fitter* ft, int enabled) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    ft->profile = enabled != 0;
    return AVB_OK;
}

int avb_fitter_groups(avb_fitter* ft, int32_t* num_groups, int32_t* joints16, int32_t* vertices16) {
    if (!ft || !num_groups) return fail(AVB_ERR_INVALID, "null argument");
    *num_groups = (int32_t)ft->group_nj.size();
    for (size_t g = 0; g < ft->group_nj.size() && g < 16; ++g) {
        if (joints16) joints16[g] = ft->group_nj[g];
        if (vertices16) vertices16[g] = ft->group_nv[g];
    }
    return AVB_OK;
}

int avb_last_flow_task_ms(avb_fitter* ft, float* ms4) {
    if (!ft || !ms4) return fail(AVB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    unsigned long long ns[4];
    CUDA_TRY(cudaMemcpy(ns, ft->d_qprof, 32, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; ++k) ms4[k] = (float)((double)ns[k] * 1e-6);
    return AVB_OK;
}

int avb_last_flow_phase_ms(avb_fitter* ft, float* ms12) {
    if (!ft || !ms12) return fail(AVB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    unsigned long long ns[16];
    CUDA_TRY(cudaMemcpy(ns, ft->d_qprof, 128, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 12; ++k) ms12[k] = (float)((double)ns[4 + k] * 1e-6);
    return AVB_OK;
}

int avb_last_kernel_ms(avb_fitter* ft, float* total_ms8, int32_t* launches8) {
    if (!ft || !total_ms8 || !launches8) return fail(AVB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    for (int k = 0; k < KC_COUNT; ++k) { total_ms8[k] = 0.f; launches8[k] = 0; }
    for (size_t i = 0; i < ft->pcls.size(); ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ft->pev[2 * i], ft->pev[2 * i + 1]));
        total_ms8[ft->pcls[i]] += ms;
        ++launches8[ft->pcls[i]];
    }
    return AVB_OK;
}

int avb_download_results(avb_fitter* ft, double* x_out, avb_stats* stats, double* cloud_out) {
    if (!ft) return fail(AVB_ERR_INVALID, "null fitter");
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const int B = ft->batch;
    const size_t nx = ft->model->nx, V = ft->model->V;
    if (x_out) CUDA_TRY(cudaMemcpyAsync(ft->h_x, ft->d_x, (size_t)B * nx * 8, cudaMemcpyDeviceToHost, st));
    if (stats) CUDA_TRY(cudaMemcpyAsync(ft->h_stats, ft->d_stats, (size_t)B * sizeof(FrameStats), cudaMemcpyDeviceToHost, st));
    if (cloud_out) CUDA_TRY(cudaMemcpyAsync(cloud_out, ft->d_cloud, (size_t)B * 3 * V * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ft->h_qctrl, ft->d_qctrl, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ft->h_qctrl[3] != 0u) {   // leave the queue usable for the next call
        cudaMemset(ft->d_qctrl, 0, 32);
        cudaMemset(ft->d_qslots, 0xFF, (size_t)ft->qcap * 8);
        return fail(AVB_ERR_CUDA, "lm_flow_kernel watchdog fired: the work queue starved (internal error)");
    }
    if (x_out) std::memcpy(x_out, ft->h_x, (size_t)B * nx * 8);
    int worst = AVB_OK;
    if (stats) {
        for (int f = 0; f < B; ++f) {
            std::memcpy(&stats[f], &ft->h_stats[f], sizeof(avb_stats));
            stats[f].num_points = (int32_t)(ft->offsets[f + 1] - ft->offsets[f]);
            if (stats[f].status != AVB_OK) worst = stats[f].status;
        }
    }
    if (worst != AVB_OK) return fail(AVB_ERR_NUMERIC, "a frame held non-finite / out-of-range input (see stats[f].status)");
    return AVB_OK;
}

int avb_fit_batch(avb_fitter* ft, int32_t batch, const double* clouds, const int32_t* labels, const int64_t* offsets,
                  double* x, const avb_options* o, avb_stats* stats, double* cloud_out) {
    if (!x) return fail(AVB_ERR_INVALID, "null parameter vector");
    int rc = avb_upload_batch(ft, batch, clouds, labels, offsets);
    if (rc != AVB_OK) return rc;
    rc = avb_fit_resident(ft, x, o);
    if (rc != AVB_OK) return rc;
    std::vector<avb_stats> tmp;
    if (!stats) {
        tmp.resize(batch);
        stats = tmp.data();
    }
    return avb_download_results(ft, x, stats, cloud_out);
}

int avb_fit(avb_fitter* ft, const double* cloud, const int32_t* labels, int32_t n, double* x, const avb_options* o,
            avb_stats* stats, double* cloud_out) {
    if (n < 0) return fail(AVB_ERR_INVALID, "negative point count");
    const int64_t offsets[2] = {0, n};
    return avb_fit_batch(ft, 1, cloud, labels, offsets, x, o, stats, cloud_out);
}

int avb_avatar_update(avb_fitter* ft, int batch, const double* x, double* cloud, double* joint_pos, double* joint_trans) {
    if (!ft || !x || batch <= 0) return fail(AVB_ERR_INVALID, "null argument");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const size_t nx = ft->model->nx, V = ft->model->V, J = ft->model->J;
    CUDA_TRY(cudaMemcpyAsync(ft->d_xdbg, x, (size_t)batch * nx * 8, cudaMemcpyHostToDevice, st));
    PoseArgs pa = pose_args(ft, ft->d_xdbg, false, nullptr);
    pa.joint_pos = ft->d_jpos;
    pa.joint_trans = ft->d_jtrans;
    CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, batch, st));
    if (cloud) CUDA_TRY(cudaMemcpyAsync(cloud, ft->d_cloud, (size_t)batch * 3 * V * 8, cudaMemcpyDeviceToHost, st));
    if (joint_pos) CUDA_TRY(cudaMemcpyAsync(joint_pos, ft->d_jpos, (size_t)batch * 3 * J * 8, cudaMemcpyDeviceToHost, st));
    if (joint_trans) CUDA_TRY(cudaMemcpyAsync(joint_trans, ft->d_jtrans, (size_t)batch * 12 * J * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AVB_OK;
}

/* -------- AvatarRenderer on the device -------- */
int avb_render_batch(avb_fitter* ft, int32_t batch, const double* x, const avb_render_desc* d, float* depth_out,
                     uint8_t* parts_out, int32_t* faces_out) {
    if (!ft || !x || !d || batch <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    if (d->width <= 0 || d->height <= 0 || (size_t)d->width * d->height >= ((size_t)1 << 31)) return fail(AVB_ERR_INVALID, "bad image size");
    if (ft->model->F > render_max_faces()) return fail(AVB_ERR_CAPACITY, "the renderer sorts at most 16384 faces per frame");
    if (ft->model->F <= 0) return fail(AVB_ERR_INVALID, "the model has no mesh (hasMesh == false): nothing to render");
    if (!depth_out && !parts_out && !faces_out) return AVB_OK;
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const size_t nx = ft->model->nx, V = ft->model->V, F = ft->model->F;
    const size_t px = (size_t)d->width * d->height, npx = px * batch;
    if (npx > ft->render_cap) {
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(ft->d_win); cudaFree(ft->d_rdepth); cudaFree(ft->d_rparts); cudaFree(ft->d_rfaces);
        ft->d_win = nullptr; ft->d_rdepth = nullptr; ft->d_rparts = nullptr; ft->d_rfaces = nullptr;
        ft->render_cap = 0;
        if (cudaMalloc(&ft->d_win, npx * 3 * 4) != cudaSuccess || cudaMalloc(&ft->d_rdepth, npx * 4) != cudaSuccess ||
            cudaMalloc(&ft->d_rparts, npx) != cudaSuccess || cudaMalloc(&ft->d_rfaces, npx * 4) != cudaSuccess) {
            cudaGetLastError();
            return fail(AVB_ERR_CUDA, "cudaMalloc of the render buffers failed");
        }
        ft->render_cap = npx;
    }
    if (!ft->d_proj) {
        CUDA_TRY(cudaMalloc(&ft->d_proj, (size_t)ft->max_batch * V * 8));
        CUDA_TRY(cudaMalloc(&ft->d_order, (size_t)ft->max_batch * F * 4));
        for (auto& e : ft->nev) CUDA_TRY(cudaEventCreate(&e));
    }
    // the posed models: Avatar::update at x (the renderer reads ava.cloud)
    CUDA_TRY(cudaMemcpyAsync(ft->d_xdbg, x, (size_t)batch * nx * 8, cudaMemcpyHostToDevice, st));
    PoseArgs pa = pose_args(ft, ft->d_xdbg, false, nullptr);
    CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, batch, st));
    CUDA_TRY(cudaMemsetAsync(ft->d_win, 0, npx * 3 * 4, st));
    RenderArgs a{};
    a.cloud = ft->d_cloud;
    a.faces = ft->dm.faces;
    a.vpart = ft->d_vpart;
    a.proj = ft->d_proj;
    a.order = ft->d_order;
    a.win_depth = depth_out ? ft->d_win : nullptr;
    a.win_parts = parts_out ? ft->d_win + npx : nullptr;
    a.win_faces = faces_out ? ft->d_win + 2 * npx : nullptr;
    a.depth_out = depth_out ? ft->d_rdepth : nullptr;
    a.parts_out = parts_out ? ft->d_rparts : nullptr;
    a.faces_out = faces_out ? ft->d_rfaces : nullptr;
    a.V = (int)V; a.F = (int)F; a.width = d->width; a.height = d->height;
    a.fx = d->fx; a.cx = d->cx; a.fy = d->fy; a.cy = d->cy;
    CUDA_TRY(launch_render(a, batch, st, ft->nev));
    if (depth_out) CUDA_TRY(cudaMemcpyAsync(depth_out, ft->d_rdepth, npx * 4, cudaMemcpyDeviceToHost, st));
    if (parts_out) CUDA_TRY(cudaMemcpyAsync(parts_out, ft->d_rparts, npx, cudaMemcpyDeviceToHost, st));
    if (faces_out) CUDA_TRY(cudaMemcpyAsync(faces_out, ft->d_rfaces, npx * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AVB_OK;
}

/* AvatarRenderer::renderLambert (AvatarRenderer.cpp:103-172) for a batch of parameter vectors */
int avb_render_lambert_batch(avb_fitter* ft, int32_t batch, const double* x, const avb_render_desc* d, uint8_t* gray_out) {
    if (!ft || !x || !d || !gray_out || batch <= 0) return fail(AVB_ERR_INVALID, "null argument or empty batch");
    if (batch > ft->max_batch) return fail(AVB_ERR_CAPACITY, "batch exceeds fitter capacity");
    if (d->width <= 0 || d->height <= 0 || (size_t)d->width * d->height >= ((size_t)1 << 31)) return fail(AVB_ERR_INVALID, "bad image size");
    if (ft->model->F > render_max_faces()) return fail(AVB_ERR_CAPACITY, "the renderer sorts at most 16384 faces per frame");
    if (ft->model->F <= 0) return fail(AVB_ERR_INVALID, "the model has no mesh (hasMesh == false): nothing to render");
    if (ft->max_valence > render_max_valence()) return fail(AVB_ERR_CAPACITY, "renderLambert handles at most 32 faces per vertex");
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    const size_t nx = ft->model->nx, V = ft->model->V, F = ft->model->F;
    const size_t px = (size_t)d->width * d->height, npx = px * batch;
    if (npx > ft->lam_cap) {
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(ft->d_win_l); cudaFree(ft->d_rlam);
        ft->d_win_l = nullptr; ft->d_rlam = nullptr; ft->lam_cap = 0;
        if (cudaMalloc(&ft->d_win_l, npx * 4) != cudaSuccess || cudaMalloc(&ft->d_rlam, npx) != cudaSuccess) {
            cudaGetLastError();
            return fail(AVB_ERR_CUDA, "cudaMalloc of the render buffers failed");
        }
        ft->lam_cap = npx;
    }
    if (!ft->d_proj) {
        CUDA_TRY(cudaMalloc(&ft->d_proj, (size_t)ft->max_batch * V * 8));
        CUDA_TRY(cudaMalloc(&ft->d_order, (size_t)ft->max_batch * F * 4));
        for (auto& e : ft->nev) CUDA_TRY(cudaEventCreate(&e));
    }
    if (!ft->d_rank_of) {
        CUDA_TRY(cudaMalloc(&ft->d_rank_of, (size_t)ft->max_batch * F * 4));
        CUDA_TRY(cudaMalloc(&ft->d_vlam, (size_t)ft->max_batch * V * 4));
    }
    CUDA_TRY(cudaMemcpyAsync(ft->d_xdbg, x, (size_t)batch * nx * 8, cudaMemcpyHostToDevice, st));
    PoseArgs pa = pose_args(ft, ft->d_xdbg, false, nullptr);
    CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, batch, st));
    CUDA_TRY(cudaMemsetAsync(ft->d_win_l, 0, npx * 4, st));
    RenderArgs a{};
    a.cloud = ft->d_cloud;
    a.faces = ft->dm.faces;
    a.vpart = ft->d_vpart;
    a.proj = ft->d_proj;
    a.order = ft->d_order;
    a.V = (int)V; a.F = (int)F; a.width = d->width; a.height = d->height;
    a.fx = d->fx; a.cx = d->cx; a.fy = d->fy; a.cy = d->cy;
    a.vf_start = ft->d_vf_start;
    a.vf_list = ft->d_vf_list;
    a.rank_of = ft->d_rank_of;
    a.vlam = ft->d_vlam;
    a.win_lambert = ft->d_win_l;
    a.lambert_out = ft->d_rlam;
    CUDA_TRY(cudaEventRecord(ft->nev[0], st));   // avb_last_render_ms after a Lambert call: {all kernels, 0, 0}
    CUDA_TRY(launch_render_lambert(a, batch, st));
    for (int k = 1; k < 4; ++k) CUDA_TRY(cudaEventRecord(ft->nev[k], st));
    CUDA_TRY(cudaMemcpyAsync(gray_out, ft->d_rlam, npx, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AVB_OK;
}

int avb_last_render_ms(avb_fitter* ft, float* ms3) {
    if (!ft || !ms3) return fail(AVB_ERR_INVALID, "null argument");
    if (!ft->nev[0]) return fail(AVB_ERR_INVALID, "no avb_render_batch yet");
    CUDA_TRY(cudaSetDevice(ft->device));
    CUDA_TRY(cudaEventSynchronize(ft->nev[3]));
    for (int k = 0; k < 3; ++k) CUDA_TRY(cudaEventElapsedTime(&ms3[k], ft->nev[k], ft->nev[k + 1]));
    return AVB_OK;
}

/* ---------------- tracking mode (BASELINE.json configs[3]) ---------------- */
int avb_track_sequence(avb_fitter* ft, int32_t T, const double* clouds, const int32_t* labels, const int64_t* offsets,
                       const double* x0, const avb_options* o, double* x_out, avb_stats* stats) {
    if (!ft || !offsets || !x0 || !x_out || T <= 0) return fail(AVB_ERR_INVALID, "null argument or empty sequence");
    int rc = check_options(ft, o);
    if (rc != AVB_OK) return rc;
    const int64_t total = offsets[T] - offsets[0], o0 = offsets[0];
    if (total < 0 || total > ft->max_points) return fail(AVB_ERR_CAPACITY, "sequence point count exceeds fitter capacity");
    if (total > 0 && (!clouds || !labels)) return fail(AVB_ERR_INVALID, "null cloud or labels");
    CUDA_TRY(cudaSetDevice(ft->device));
    cudaStream_t st = ft->stream;
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!ft->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ft->copy_stream, cudaStreamNonBlocking));
    while ((int)ft->copied.size() < T) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ft->copied.push_back(e);
    }
    const size_t nx = ft->model->nx, V = ft->model->V;
    if (ft->seq_cap < T) {
        dev_release(ft, &ft->d_xseq); dev_release(ft, &ft->d_stats_seq);
        pin_release(ft, &ft->h_xseq); pin_release(ft, &ft->h_stats_seq);
        ft->seq_cap = 0;
        rc = dev_alloc(ft, &ft->d_xseq, (size_t)T * nx);
        if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_stats_seq, (size_t)T);
        if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_xseq, (size_t)T * nx);
        if (rc == AVB_OK) rc = pin_alloc(ft, &ft->h_stats_seq, (size_t)T);
        if (rc != AVB_OK) return rc;
        ft->seq_cap = T;
    }
    // per-frame NN chunk schedule; every frame is "frame 0" of a batch of one
    std::vector<int> c0(T + 1, 0), q0(T + 1, 0);
    int nc = 0, qb = 0;
    {   // a long sequence of small frames needs more chunk slots than total points / 512: size the tables from the frames
        size_t need = 0;
        for (int t = 0; t < T; ++t) {
            const int64_t n = offsets[t + 1] - offsets[t];
            if (n < 0 || n > ((int64_t)1 << 21)) return fail(AVB_ERR_INVALID, "bad offsets");
            const int64_t target = std::max<int64_t>(512, (n / (4 * (int64_t)ft->num_sms) + 511) / 512 * 512);
            need += (size_t)((n + target - 1) / target);
        }
        rc = ensure_chunk_tables(ft, need);
        if (rc != AVB_OK) return rc;
    }
    for (int t = 0; t < T; ++t) {
        const int64_t n = offsets[t + 1] - offsets[t];
        c0[t] = nc;
        q0[t] = qb;
        if (n > 0) {
            const int64_t target = std::max<int64_t>(512, (n / (4 * (int64_t)ft->num_sms) + 511) / 512 * 512);
            for (int64_t b = 0; b < n; b += target) {
                if (nc >= ft->max_chunks) return fail(AVB_ERR_CAPACITY, "too many NN chunks for this fitter (raise max_total_points)");
                ft->h_chunk_frame[nc] = 0;
                ft->h_chunk_begin[nc] = offsets[t] - o0 + b;
                ft->h_chunk_count[nc] = (int)std::min<int64_t>(target, n - b);
                ft->h_chunk_qblock[nc] = (int)(b / kQBlock);   // relative to the frame's first |d|^2 block
                ++nc;
            }
            qb += (int)((n + kQBlock - 1) / kQBlock);
        }
    }
    c0[T] = nc;
    q0[T] = qb;
    // every frame is "frame 0" of a batch of one and its |d|^2 partials start at d_qpart[0]: the largest frame must fit
    for (int t = 0; t < T; ++t)
        if (q0[t + 1] - q0[t] > ft->max_qblocks) return fail(AVB_ERR_CAPACITY, "too many |d|^2 blocks");
    if (nc > 0) {
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_frame, ft->h_chunk_frame, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_begin, ft->h_chunk_begin, (size_t)nc * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_count, ft->h_chunk_count, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_chunk_qblock, ft->h_chunk_qblock, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
    }
    std::memcpy(ft->h_x, x0, nx * 8);
    CUDA_TRY(cudaMemcpyAsync(ft->d_x, ft->h_x, nx * 8, cudaMemcpyHostToDevice, st));
    ft->data_is_f32 = false;
    // uploads run ahead on the copy stream, one event per frame
    for (int t = 0; t < T; ++t) {
        const int64_t n = offsets[t + 1] - offsets[t], b = offsets[t] - o0;
        if (n > 0) {
            CUDA_TRY(cudaMemcpyAsync(ft->d_data + 3 * b, clouds + 3 * offsets[t], (size_t)n * 24, cudaMemcpyHostToDevice, ft->copy_stream));
            CUDA_TRY(cudaMemcpyAsync(ft->d_labels + b, labels + offsets[t], (size_t)n * 4, cudaMemcpyHostToDevice, ft->copy_stream));
        }
        CUDA_TRY(cudaEventRecord(ft->copied[t], ft->copy_stream));
    }
    ft->batch = 1;
    ft->launches = 0;
    ft->last_icp = 0;
    const int all_chunks = ft->num_chunks;
    // the |d|^2 partials of frame t live at d_qpart[0 ..): frame_qblock = {0, nblocks_t}
    for (int t = 0; t < T; ++t) {
        CUDA_TRY(cudaStreamWaitEvent(st, ft->copied[t], 0));
        const int fq[2] = {0, q0[t + 1] - q0[t]};
        ft->h_frame_qblock[0] = fq[0];
        ft->h_frame_qblock[1] = fq[1];
        CUDA_TRY(cudaMemcpyAsync(ft->d_frame_qblock, ft->h_frame_qblock, 8, cudaMemcpyHostToDevice, st));
        // h_frame_qblock is rewritten next iteration: the copy above must have been consumed (tiny; sync the copy only)
        cudaEvent_t ev = ft->ev[6];
        CUDA_TRY(cudaEventRecord(ev, st));
        for (int icp = 0; icp < o->icp_iters; ++icp) {
            // chunk views of this frame
            int* sv_frame = ft->d_chunk_frame; long long* sv_begin = ft->d_chunk_begin;
            int* sv_count = ft->d_chunk_count; int* sv_qb = ft->d_chunk_qblock;
            ft->d_chunk_frame += c0[t]; ft->d_chunk_begin += c0[t]; ft->d_chunk_count += c0[t]; ft->d_chunk_qblock += c0[t];
            ft->num_chunks = c0[t + 1] - c0[t];
            rc = enqueue_correspond(ft, ft->d_x, o, nullptr);
            ft->d_chunk_frame = sv_frame; ft->d_chunk_begin = sv_begin; ft->d_chunk_count = sv_count; ft->d_chunk_qblock = sv_qb;
            if (rc != AVB_OK) { ft->num_chunks = all_chunks; return rc; }
            LmBuf la = lm_buf(ft, ft->d_x, o);
            rc = enqueue_solve(ft, la, o, 1 + o->max_iters_per_icp);
            if (rc != AVB_OK) { ft->num_chunks = all_chunks; return rc; }
        }
        CUDA_TRY(cudaMemcpyAsync(ft->d_xseq + (size_t)t * nx, ft->d_x, nx * 8, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(ft->d_stats_seq + t, ft->d_stats, sizeof(FrameStats), cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaEventSynchronize(ev));
    }
    ft->num_chunks = all_chunks;
    // trailing ava.update() of the last frame so that AVB_TAP_CLOUD / avb_download_results see it
    PoseArgs pa = pose_args(ft, ft->d_x, false, o);
    CUDA_TRY(launch_pose_visibility(ft->dm, ft->dp, pa, 1, st));
    CUDA_TRY(cudaMemcpyAsync(ft->h_xseq, ft->d_xseq, (size_t)T * nx * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ft->h_stats_seq, ft->d_stats_seq, (size_t)T * sizeof(FrameStats), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ft->h_qctrl, ft->d_qctrl, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ft->h_qctrl[3] != 0u) {   // a starved work queue must not return stale parameters as AVB_OK
        cudaMemset(ft->d_qctrl, 0, 32);
        cudaMemset(ft->d_qslots, 0xFF, (size_t)ft->qcap * 8);
        return fail(AVB_ERR_CUDA, "lm_flow_kernel watchdog fired: the work queue starved (internal error)");
    }
    std::memcpy(x_out, ft->h_xseq, (size_t)T * nx * 8);
    ft->offsets.assign(2, 0);
    ft->offsets[1] = offsets[T] - offsets[T - 1];
    ft->total_points = ft->offsets[1];
    (void)V;
    int worst = AVB_OK;
    for (int t = 0; t < T; ++t) {
        if (stats) {
            std::memcpy(&stats[t], &ft->h_stats_seq[t], sizeof(avb_stats));
            stats[t].num_points = (int32_t)(offsets[t + 1] - offsets[t]);
        }
        if (ft->h_stats_seq[t].status != AVB_OK) worst = ft->h_stats_seq[t].status;
    }
    if (worst != AVB_OK) return fail(AVB_ERR_NUMERIC, "a frame held non-finite / out-of-range input (see stats[t].status)");
    return AVB_OK;
}

/* ---------------- parity taps ---------------- */
int avb_debug_correspond(avb_fitter* ft, const double* x_in, const avb_options* o) {
    if (!ft || !x_in) return fail(AVB_ERR_INVALID, "null argument");
    if (ft->batch <= 0) return fail(AVB_ERR_INVALID, "no batch uploaded");
    int rc = check_options(ft, o);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t nx = ft->model->nx;
    CUDA_TRY(cudaMemcpyAsync(ft->d_xdbg, x_in, (size_t)ft->batch * nx * 8, cudaMemcpyHostToDevice, ft->stream));
    ft->launches = 0;
    rc = enqueue_correspond(ft, ft->d_xdbg, o, nullptr);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    return AVB_OK;
}

int avb_debug_read(avb_fitter* ft, int what, void* out, uint64_t bytes) {
    if (!ft || !out) return fail(AVB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t B = ft->batch, V = ft->model->V;
    const void* src = nullptr;
    size_t need = 0;
    switch (what) {
        case AVB_TAP_VISIBLE: src = ft->d_vis; need = B * V; break;
        case AVB_TAP_NN: src = ft->d_nn; need = (size_t)ft->total_points * 4; break;
        case AVB_TAP_CLOUD: src = ft->d_cloud; need = B * 3 * V * 8; break;
        case AVB_TAP_COUNT: src = ft->d_cnt; need = B * V * 4; break;
        case AVB_TAP_SUM: need = B * 3 * V * 8; break;
        default: return fail(AVB_ERR_INVALID, "unknown tap");
    }
    if (bytes < need) return fail(AVB_ERR_INVALID, "tap buffer too small");
    CUDA_TRY(cudaStreamSynchronize(ft->stream));
    if (what == AVB_TAP_SUM) {
        std::vector<long long> tmp(B * 3 * V);
        CUDA_TRY(cudaMemcpy(tmp.data(), ft->d_sum, need, cudaMemcpyDeviceToHost));
        double* o = static_cast<double*>(out);
        for (size_t i = 0; i < tmp.size(); ++i) o[i] = (double)tmp[i] * kFixInv;
        return AVB_OK;
    }
    CUDA_TRY(cudaMemcpy(out, src, need, cudaMemcpyDeviceToHost));
    return AVB_OK;
}

int avb_debug_evaluate(avb_fitter* ft, const double* x_in, const avb_options* o, double* cost, double* grad, double* H) {
    if (!ft || !x_in || !cost || !grad || !H) return fail(AVB_ERR_INVALID, "null argument");
    if (ft->batch <= 0) return fail(AVB_ERR_INVALID, "no batch uploaded");
    int rc = check_options(ft, o);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaSetDevice(ft->device));
    const size_t B = ft->batch, nx = ft->model->nx, P = ft->model->P;
    if (!ft->d_dump_cost) {
        rc = dev_alloc(ft, &ft->d_dump_cost, (size_t)ft->max_batch);
        if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_dump_grad, (size_t)ft->max_batch * P);
        if (rc == AVB_OK) rc = dev_alloc(ft, &ft->d_dump_H, (size_t)ft->max_batch * P * P);
        if (rc != AVB_OK) return rc;
    }
    cudaStream_t st = ft->stream;
    CUDA_TRY(cudaMemcpyAsync(ft->d_xdbg, x_in, B * nx * 8, cudaMemcpyHostToDevice, st));
    LmBuf la = lm_buf(ft, ft->d_xdbg, o);
    la.max_iters = 0;
    la.dump_cost = ft->d_dump_cost;
    la.dump_grad = ft->d_dump_grad;
    la.dump_H = ft->d_dump_H;
    rc = enqueue_solve(ft, la, o, 1);
    if (rc != AVB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(cost, ft->d_dump_cost, B * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(grad, ft->d_dump_grad, B * P * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(H, ft->d_dump_H, B * P * P * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return AVB_OK;
}

}  // extern "C"
