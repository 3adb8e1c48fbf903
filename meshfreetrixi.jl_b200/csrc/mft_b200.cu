// mft_b200.cu -- C ABI (include/mft_b200.h) + host orchestration of the sm_100a rhs! path.
// There is no CPU fallback anywhere in this file: every compute entry point needs a CUDA device.
#include "../../include/mft_b200.h"
#include "mft_kernels.cuh"
#include "mft_fused_kernels.cuh"
#include "mft_tile_kernels.cuh"
#include "mft_limiter_kernels.cuh"
#include "mft_igr_kernels.cuh"
#include "mft_aux_kernels.cuh"
#include "mft_nccl.h"

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: a no-op unless a profiler injects itself

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

using namespace mft;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(x)                                                                                               \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess) return fail(MFT_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)
#define CHECK(x)             \
    do {                     \
        int r_ = (x);        \
        if (r_ != 0) return r_; \
    } while (0)

// ------------------------------------------------------------------------------------------------------
// host-side sparse helpers (setup time)
// ------------------------------------------------------------------------------------------------------
struct HostCsc {
    bool set = false;
    std::vector<int64_t> colptr, rowval;  // 0-based internally
    std::vector<double> nz;
};

struct Csr2 {  // rows with paired weights, entries of a row in summation order
    int64_t nrows = 0;
    std::vector<int64_t> ptr;
    std::vector<int32_t> col;
    std::vector<double> wx, wy;
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    int64_t n = 0;
    int alloc(int64_t count)
    {
        release();
        n = count;
        if (count == 0) return 0;
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (size_t)count);
        if (e != cudaSuccess) return fail(MFT_ECUDA, "cudaMalloc(%lld bytes): %s", (long long)(sizeof(T) * count), cudaGetErrorString(e));
        return 0;
    }
    int upload(const std::vector<T> &h)
    {
        CHECK(alloc((int64_t)h.size()));
        if (!h.empty()) CU(cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct DevEll {
    DevBuf<unsigned char> blob;
    DevBuf<int> off;
    int nslices = 0;
    int colb = 0;             // bytes per column: kColBytes2 (paired) or kColBytes1
    int maxw = 0;             // widest slice
    int64_t ncols_total = 0;  // sum of slice widths
    int64_t nnz = 0;
    EllBlob view() const { return EllBlob{blob.p, off.p}; }
    void release()
    {
        blob.release();
        off.release();
    }
};

// union tiles with R rows per thread (mask-compressed weights), see mft_tile_kernels.cuh
struct DevTileR {
    DevBuf<unsigned char> blob;
    DevBuf<long long> boff;
    DevBuf<int> wl, uoff, ulist;
    DevBuf<unsigned short> uslot;
    int R = 0, nslices = 0, ntiles = 0, maxW = 0, maxL = 0, sstride = 0, ncopy = 1;
    int64_t nnz = 0, nunion = 0, nsteps = 0;
    // several GPUs, fused step: block -> tile, interior tiles first; the last ntiles - n_free blocks are band tiles
    DevBuf<int> order;
    int n_free = 0;
    bool ready() const { return blob.p != nullptr; }
    void release()
    {
        order.release();
        blob.release();
        boff.release();
        wl.release();
        uoff.release();
        ulist.release();
        uslot.release();
    }
};

struct BcGroup {
    int kind;
    int64_t nb;
    DevBuf<int> idx;
    DevBuf<double> normals, values;
    // time-dependent Dirichlet data inside a device-resident step: tables for the two stage times of the step
    // (slot 0: t + dt, slot 1: t + dt/2), copied over `values` in front of the rhs! evaluated at that time
    DevBuf<double> stage_values[2];
    bool stage_set[2] = {false, false};
};

struct Source {
    int kind;
    double gamma = 0, c_rv = 1, c_uw = 1, dx_avg = 0;
    int polydeg = 4;
    HostCsc hv_host;
    DevEll hv;
    // SourceIGR (IGR.jl): alpha, CG iteration cap, forward operator as plain sliced ELL, CG vectors, device-side CG scalars
    double igr_alpha = 1.0;
    int igr_maxiter = 20;
    DevEll igr_op;
    DevBuf<double> igr_vec;  // rho_inv, b, r, c (n each), x, p (n_tot + 1 each), t (2 (n_tot + 1))
    DevBuf<double> igr_partial;
    DevBuf<unsigned int> igr_ticket;
    DevBuf<mft_igr::IgrScalars> igr_scalars;
    mft_igr::IgrArgs igr_args;
};

struct KTimer {
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<int> cls;
};

struct mft_ctx {
    int device = 0;
    int64_t n_local = 0, n_halo = 0, n_tot = 0;
    int V = 0, ndims = 2, k = 0;
    int eq = -1;
    double eqp[2] = {0, 0};
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // options
    int exact = 1, mean_div_vn = 1, max_lex = 1, diagnostics = 0, stage_w = 1, stage_w_b = 0, pf_dist = 0, refine_order = 0, kfix_ok = 0;
    std::vector<const void *> smem_configured;
    // ordering
    bool have_perm = false;
    std::vector<int32_t> perm, iperm;  // device->caller, caller->device (0-based)
    std::vector<int64_t> keys;         // caller index -> summation rank
    DevBuf<int> d_perm;
    // operators
    HostCsc host_ops[2];
    Csr2 host_ell;  // from mft_set_operator_ell
    bool have_ell_input = false;
    DevEll fwd, fwd_pair, tra, tra_pair;
    DevTileR fwd_tiler, tra_tiler;
    // Zhang-Shu limiter: kNN table in list order, scratch, stage-limiter configuration
    std::vector<int32_t> host_nbr;  // caller numbering, n_local x k row-major
    DevBuf<int> zs_nbr;
    DevBuf<double> zs_tmp;
    DevBuf<unsigned char> zs_flag;
    std::vector<double> stage_lim_thresholds;
    std::vector<int> stage_lim_variables;
    // fused step (mft_fused_kernels.cuh): per-row side table (boundary entry, halo routes), miss counter of the one-pass norms
    int fused_step = 1;
    int pdl = 1;                           // MFT_OPT_PDL: programmatic dependent launch between the kernels of a fused stage
    int layout_device = 1;                 // MFT_OPT_LAYOUT_DEVICE: R = 1 union tiles are laid out by the device builder
    bool fused_active = false;             // the launches being issued belong to the fused step
    bool pdl_next = false;                 // the next fused kernel follows another kernel of the fused step directly
    DevBuf<int> row_aux;
    DevBuf<RowAux> row_aux_tab;
    DevBuf<int> route_peer;
    DevBuf<long long> route_dst;
    DevBuf<unsigned long long> norm_miss;
    unsigned long long norm_miss_seen = 0;   // value of the counter at the last check (mft_synchronize / mft_download_state)
    bool fused_used = false;
    DevBuf<P2PPeers> peers_dev_buf;        // device copy of peers_dev for kernels that take a pointer
    std::vector<int> bc_idx_host;          // merged boundary table: device rows
    std::vector<int> send_rows_host, send_peer_host;
    std::vector<long long> send_dst_host;
    int tile_rows_a = 1, tile_rows_b = 1;  // MFT_OPT_TILE_ROWS: rows per thread of the union-tile kernels (1, 2 or 4)
    int stage_force = 0;
    int tile = 31;  // MFT_OPT_TILE: bit 0 pass A, bit 1 pass B (Euler 2-D only), bit 2 bank-coloured slots, bit 3 two copies,
                    // bit 4 tuned second copy (default: profiles/README.md r2)
    int pair_rows = 1;
    int two_phase = 1;
    // bcs, sources
    std::vector<BcGroup *> bcs;
    std::vector<Source *> srcs;
    bool finalized = false;
    // device state (AoS)
    DevBuf<double> u, du, uprev, g, approx_du, stage_soa;
    DevBuf<double> utilde, kfsal, u_save;  // SSPRK43: error vector, f(u_n) kept for a rejected step, u_n incl. halo
    bool step_pending = false;
    std::vector<DevBuf<double>> hist;
    std::vector<double> time_history, time_weights;
    int hist_head = 0, nslots = 0;
    int64_t success_iter = 0;
    // diagnostics
    DevBuf<double> eps, eps_uw, eps_rv, eps_c, residual;
    // reductions
    DevBuf<double> partial, stats;  // stats: sum[V], mean[V], norms[V]
    DevBuf<unsigned int> ticket;
    // merged boundary table (all groups, one launch) when no point is in two groups
    // rows of u that rhs! can change (boundary points, halo tail): the only part of u a host caller gets back
    std::vector<int64_t> touched_caller;
    DevBuf<int> touched_dev;
    DevBuf<double> touched_buf;
    double *touched_host = nullptr;
    std::vector<std::vector<int64_t>> bc_caller_rows;
    bool bc_merged = false;
    int64_t bc_total = 0;
    std::vector<int64_t> bc_group_off;
    DevBuf<int> bc_kind, bc_idx;
    DevBuf<double> bc_normals, bc_values;
    int red_blocks = 0;
    bool have_fsal = false;
    int64_t launches = 0;
    // captured SSPRK steps (CUDA graphs), keyed by the launch parameters that are baked into kernel arguments
    struct StepGraph {
        double dt;
        int si_zero, with_first_rhs, uses;
        int64_t nlaunch;
        cudaGraphExec_t exec;
    };
    std::vector<StepGraph> graphs;
    int use_graphs = 1;
    // per-class timing
    bool timing = false;
    KTimer kt;
    // multi-GPU
    NcclApi *nccl = nullptr;
    void *comm = nullptr;
    int nranks = 1, rank = 0;
    int64_t n_global = 0;
    std::vector<int> peers;
    std::vector<int64_t> send_off, recv_off;  // prefix offsets (points)
    DevBuf<int> send_rows;
    DevBuf<double> send_buf;
    DevBuf<double> gather_buf;  // nranks * 2V doubles
    int64_t n_send = 0;
    // peer-memory exchange (CUDA IPC over NVLink)
    bool p2p = false;
    P2PPeers peers_dev{};
    DevBuf<unsigned char> p2p_window, p2p_local;
    DevBuf<int> send_peer;
    DevBuf<long long> send_dst;
    std::vector<void *> ipc_opened;
};

// ------------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------------
extern "C" const char *mft_last_error(void) { return g_err.c_str(); }
extern "C" int mft_version(void) { return 120; }  // 1.1: setup pipeline, limiter, IGR source, non-finite check; 1.2: stage-time Dirichlet tables
extern "C" int mft_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// NVTX ranges carrying the labels of the reference's TimerOutputs sections (`@trixi_timeit timer() "..."`:
// rbfsolver.jl:392,400-425, parallel_rbfsolver.jl:96-132, history.jl:69), so an nsys / ncu timeline of the GPU path reads
// like the reference's timer table.  Host-side only: nothing is recorded during CUDA-graph replay (one range per replay).
struct NvtxRange {
    explicit NvtxRange(const char *label) { nvtxRangePushA(label); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct ScopedTimer {
    mft_ctx *c;
    bool on;
    ScopedTimer(mft_ctx *ctx, int cls) : c(ctx), on(ctx->timing)
    {
        if (on) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            c->kt.ev.push_back(a);
            c->kt.ev.push_back(b);
            c->kt.cls.push_back(cls);
            cudaEventRecord(a, c->stream);
        }
    }
    ~ScopedTimer()
    {
        if (on) cudaEventRecord(c->kt.ev.back(), c->stream);
    }
};

#define LAUNCH_CHECK()                                                                              \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ != cudaSuccess) return fail(MFT_ECUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

// Kernel launch, optionally as a programmatic dependent of the previous kernel on the stream (PDL): the kernel may begin while
// its predecessor drains; it calls pdl_wait() before it reads the predecessor's output.  Captured into CUDA graphs as a
// programmatic edge.
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------------
extern "C" int mft_ctx_create(mft_ctx **out, int device, int64_t n_local, int64_t n_halo, int nvars, int ndims, int k)
{
    if (!out) return fail(MFT_EINVAL, "mft_ctx_create: out is NULL");
    *out = nullptr;
    if (n_local <= 0 || n_halo < 0) return fail(MFT_EINVAL, "mft_ctx_create: bad sizes n_local=%lld n_halo=%lld", (long long)n_local, (long long)n_halo);
    if (n_local + n_halo > 2000000000LL) return fail(MFT_EINVAL, "mft_ctx_create: more than 2^31 points per device is not supported");
    if (nvars != 1 && nvars != 4) return fail(MFT_ENOTSUP, "mft_ctx_create: nvars must be 1 (advection) or 4 (Euler 2-D), got %d", nvars);
    if (ndims != 2) return fail(MFT_ENOTSUP, "mft_ctx_create: only 2-D point clouds are supported, got ndims=%d", ndims);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MFT_ENODEVICE, "mft_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(MFT_EINVAL, "mft_ctx_create: device %d out of range [0,%d)", device, ndev);
    CU(cudaSetDevice(device));
    mft_ctx *c = new mft_ctx();
    c->device = device;
    c->n_local = n_local;
    c->n_halo = n_halo;
    c->n_tot = n_local + n_halo;
    c->V = nvars;
    c->ndims = ndims;
    c->k = k;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming));
    CU(cudaEventCreate(&c->ev_t0));
    CU(cudaEventCreate(&c->ev_t1));
    const int64_t len = c->n_tot * nvars;
    CHECK(c->u.alloc(len + nvars));  // + the dummy record gathered by padding entries
    CHECK(c->du.alloc(len));
    CHECK(c->stage_soa.alloc(len));
    CU(cudaMemset(c->u.p, 0, sizeof(double) * (len + nvars)));
    CU(cudaMemset(c->du.p, 0, sizeof(double) * len));
    if (nvars == 4) {  // a finite Euler state so that the flux of the dummy record is finite (its weight is 0)
        const double dummy[4] = {1.0, 0.0, 0.0, 1.0};
        CU(cudaMemcpy(c->u.p + len, dummy, sizeof dummy, cudaMemcpyHostToDevice));
    }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    // one wave of the fused stage kernel (MFT_STAGE_OCC blocks of 256 threads per SM); the separate reduction kernels use the
    // same grid, so both paths sum in the same order
    c->red_blocks = prop.multiProcessorCount * MFT_STAGE_OCC;
    // per-block partials: V sums, or one norm record (fused step) + one record per group of kStageGroup blocks
    const int ngroups = (c->red_blocks + kStageGroup - 1) / kStageGroup;
    CHECK(c->partial.alloc((int64_t)(c->red_blocks + ngroups) * kRecDoubles));
    CHECK(c->stats.alloc(24));  // sum | mean | norms | SSPRK43 error sum | [16..19] raw norms of the fused step
    CHECK(c->ticket.alloc(8 + ngroups));   // [8..]: group tickets of the fused stage kernel
    CU(cudaMemset(c->ticket.p, 0, (8 + ngroups) * sizeof(unsigned int)));
    CHECK(c->p2p_local.alloc((int64_t)sizeof(P2PLocal)));   // norms epoch / norm_ready of the fused step (also on one GPU)
    CU(cudaMemset(c->p2p_local.p, 0, sizeof(P2PLocal)));
    CHECK(c->norm_miss.alloc(1));
    CU(cudaMemset(c->norm_miss.p, 0, sizeof(unsigned long long)));
    c->pf_dist = prop.multiProcessorCount * 8;
    CU(cudaMemset(c->stats.p, 0, sizeof(double) * 24));
    *out = c;
    return MFT_OK;
}

extern "C" int mft_ctx_destroy(mft_ctx *c)
{
    if (!c) return MFT_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->comm && c->nccl) c->nccl->commDestroy(c->comm);
    for (auto *b : c->bcs) {
        b->idx.release();
        b->normals.release();
        b->values.release();
        b->stage_values[0].release();
        b->stage_values[1].release();
        delete b;
    }
    for (auto *s : c->srcs) {
        s->hv.release();
        s->igr_op.release();
        s->igr_vec.release();
        s->igr_partial.release();
        s->igr_ticket.release();
        s->igr_scalars.release();
        delete s;
    }
    c->fwd.release();
    c->tra.release();
    c->tra_pair.release();
    c->fwd_pair.release();
    c->fwd_tiler.release();
    c->tra_tiler.release();
    c->zs_nbr.release();
    c->zs_tmp.release();
    c->zs_flag.release();
    for (auto &h : c->hist) h.release();
    DevBuf<double> *bufs[] = {&c->utilde, &c->kfsal, &c->u_save, &c->u, &c->du, &c->uprev, &c->g, &c->approx_du, &c->stage_soa, &c->eps, &c->eps_uw,
                              &c->eps_rv, &c->eps_c, &c->residual, &c->partial, &c->stats, &c->send_buf, &c->gather_buf};
    for (auto *b : bufs) b->release();
    if (c->touched_host) cudaFreeHost(c->touched_host);
    c->touched_dev.release();
    c->touched_buf.release();
    for (void *q : c->ipc_opened) cudaIpcCloseMemHandle(q);
    c->p2p_window.release();
    c->p2p_local.release();
    c->send_peer.release();
    c->send_dst.release();
    c->row_aux.release();
    c->row_aux_tab.release();
    c->route_peer.release();
    c->route_dst.release();
    c->norm_miss.release();
    c->peers_dev_buf.release();
    c->d_perm.release();
    c->send_rows.release();
    c->ticket.release();
    c->bc_kind.release();
    c->bc_idx.release();
    c->bc_normals.release();
    c->bc_values.release();
    for (auto &g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto ev : c->kt.ev) cudaEventDestroy(ev);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->ev_c) cudaEventDestroy(c->ev_c);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    delete c;
    return MFT_OK;
}

#define NEED_CTX(c)                                               \
    do {                                                          \
        if (!(c)) return fail(MFT_EINVAL, "%s: ctx is NULL", __func__); \
        CU(cudaSetDevice((c)->device));                           \
    } while (0)

extern "C" int mft_set_equation(mft_ctx *c, int kind, const double *params, int nparams)
{
    NEED_CTX(c);
    if (kind == MFT_EQ_EULER2D) {
        if (c->V != 4) return fail(MFT_EINVAL, "mft_set_equation: Euler 2-D needs nvars=4");
        if (nparams < 1 || !params) return fail(MFT_EINVAL, "mft_set_equation: Euler 2-D needs gamma");
        c->eqp[0] = params[0];
    } else if (kind == MFT_EQ_ADVECTION2D) {
        if (c->V != 1) return fail(MFT_EINVAL, "mft_set_equation: advection needs nvars=1");
        if (nparams < 2 || !params) return fail(MFT_EINVAL, "mft_set_equation: advection needs a1, a2");
        c->eqp[0] = params[0];
        c->eqp[1] = params[1];
    } else {
        return fail(MFT_ENOTSUP, "mft_set_equation: unknown equation kind %d", kind);
    }
    c->eq = kind;
    return MFT_OK;
}

extern "C" int mft_set_option(mft_ctx *c, int option, double value)
{
    NEED_CTX(c);
    switch (option) {
    case MFT_OPT_EXACT_ORDER: c->exact = value != 0; break;
    case MFT_OPT_MEAN_DIVISOR_VN: c->mean_div_vn = value != 0; break;
    case MFT_OPT_MAX_LEXICOGRAPHIC: c->max_lex = value != 0; break;
    case MFT_OPT_DIAGNOSTICS: c->diagnostics = value != 0; break;
    case MFT_OPT_CUDA_GRAPH: c->use_graphs = (int)value; break;
    case MFT_OPT_STAGE_WEIGHTS: c->stage_w = ((int)value & 1) != 0; c->stage_w_b = ((int)value & 2) != 0; c->two_phase = ((int)value & 4) != 0; c->stage_force = ((int)value & 8) != 0; break;
    case MFT_OPT_PREFETCH_DISTANCE: c->pf_dist = (int)value; break;
    case MFT_OPT_REFINE_ORDER: c->refine_order = value != 0; break;
    case MFT_OPT_SINGLE_SWEEP_EXACT: c->kfix_ok = value != 0; break;
    case MFT_OPT_PAIR_ROWS: c->pair_rows = (int)value; break;
    case MFT_OPT_TILE: c->tile = (int)value; break;
    case MFT_OPT_LAYOUT_DEVICE: c->layout_device = value != 0; break;
    case MFT_OPT_PDL:
        c->pdl = value != 0;
        for (auto &g : c->graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        c->graphs.clear();
        break;
    case MFT_OPT_FUSED_STEP:
        c->fused_step = value != 0;
        for (auto &g : c->graphs)   // captured steps bake the launch sequence in: start over
            if (g.exec) cudaGraphExecDestroy(g.exec);
        c->graphs.clear();
        break;
    case MFT_OPT_TILE_ROWS: {
        const int a = (int)value % 10, b = ((int)value / 10) % 10;
        if ((a != 1 && a != 2 && a != 4) || (b != 1 && b != 2 && b != 4)) return fail(MFT_EINVAL, "MFT_OPT_TILE_ROWS: digits must be 1, 2 or 4");
        c->tile_rows_a = a;
        c->tile_rows_b = b;
        break;
    }
    default: return fail(MFT_EINVAL, "mft_set_option: unknown option %d", option);
    }
    return MFT_OK;
}

extern "C" int mft_set_permutation(mft_ctx *c, const int64_t *perm1)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_set_permutation: must be called before the first compute call");
    if (!perm1) return fail(MFT_EINVAL, "mft_set_permutation: perm is NULL");
    const int64_t n = c->n_tot;
    std::vector<int32_t> perm(n), iperm(n, -1);
    for (int64_t d = 0; d < n; ++d) {
        const int64_t p = perm1[d] - 1;
        if (p < 0 || p >= n) return fail(MFT_EINVAL, "mft_set_permutation: entry %lld out of range", (long long)perm1[d]);
        if (iperm[p] != -1) return fail(MFT_EINVAL, "mft_set_permutation: duplicate entry %lld", (long long)perm1[d]);
        if ((d < c->n_local) != (p < c->n_local)) return fail(MFT_EINVAL, "mft_set_permutation: owned and halo points must not mix");
        perm[d] = (int32_t)p;
        iperm[p] = (int32_t)d;
    }
    c->perm.swap(perm);
    c->iperm.swap(iperm);
    c->have_perm = true;
    return MFT_OK;
}

extern "C" int mft_set_order_keys(mft_ctx *c, const int64_t *keys)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_set_order_keys: must be called before the first compute call");
    if (!keys) return fail(MFT_EINVAL, "mft_set_order_keys: keys is NULL");
    c->keys.assign(keys, keys + c->n_tot);
    return MFT_OK;
}

static int copy_csc(mft_ctx *c, HostCsc &dst, const int64_t *colptr, const int64_t *rowval, const double *nzval, const char *who)
{
    if (!colptr || !rowval || !nzval) return fail(MFT_EINVAL, "%s: NULL CSC array", who);
    const int64_t n = c->n_tot;
    if (colptr[0] != 1) return fail(MFT_EINVAL, "%s: colptr[0] must be 1 (Julia 1-based), got %lld", who, (long long)colptr[0]);
    const int64_t nnz = colptr[n] - 1;
    if (nnz < 0) return fail(MFT_EINVAL, "%s: negative nnz", who);
    dst.colptr.resize(n + 1);
    for (int64_t i = 0; i <= n; ++i) {
        dst.colptr[i] = colptr[i] - 1;
        if (i > 0 && dst.colptr[i] < dst.colptr[i - 1]) return fail(MFT_EINVAL, "%s: colptr not monotone at %lld", who, (long long)i);
    }
    dst.rowval.resize(nnz);
    for (int64_t p = 0; p < nnz; ++p) {
        const int64_t r = rowval[p] - 1;
        if (r < 0 || r >= n) return fail(MFT_EINVAL, "%s: rowval[%lld]=%lld out of range", who, (long long)p, (long long)rowval[p]);
        dst.rowval[p] = r;
    }
    dst.nz.assign(nzval, nzval + nnz);
    dst.set = true;
    return MFT_OK;
}

extern "C" int mft_set_operator_csc(mft_ctx *c, int slot, const int64_t *colptr, const int64_t *rowval, const double *nzval)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_set_operator_csc: operators are immutable after the first compute call");
    if (slot != MFT_OP_DX && slot != MFT_OP_DY) return fail(MFT_EINVAL, "mft_set_operator_csc: slot must be MFT_OP_DX or MFT_OP_DY");
    return copy_csc(c, c->host_ops[slot], colptr, rowval, nzval, "mft_set_operator_csc");
}

extern "C" int mft_set_operator_ell(mft_ctx *c, const int64_t *nbr1, const double *wx, const double *wy)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_set_operator_ell: operators are immutable after the first compute call");
    if (!nbr1 || !wx || !wy) return fail(MFT_EINVAL, "mft_set_operator_ell: NULL array");
    if (c->k <= 0) return fail(MFT_EINVAL, "mft_set_operator_ell: ctx was created with k=%d", c->k);
    Csr2 &A = c->host_ell;
    const int64_t n = c->n_local, k = c->k;
    A.nrows = n;
    A.ptr.resize(n + 1);
    A.col.resize(n * k);
    A.wx.assign(wx, wx + n * k);
    A.wy.assign(wy, wy + n * k);
    for (int64_t i = 0; i <= n; ++i) A.ptr[i] = i * k;
    for (int64_t p = 0; p < n * k; ++p) {
        const int64_t j = nbr1[p] - 1;
        if (j < 0 || j >= c->n_tot) return fail(MFT_EINVAL, "mft_set_operator_ell: neighbour %lld out of range", (long long)nbr1[p]);
        A.col[p] = (int32_t)j;
    }
    c->have_ell_input = true;
    return MFT_OK;
}

extern "C" int mft_add_boundary(mft_ctx *c, int kind, int64_t nb, const int64_t *idx1, const double *normals, const double *values)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_add_boundary: must be called before the first compute call");
    if (kind < 0 || kind > 2) return fail(MFT_EINVAL, "mft_add_boundary: unknown kind %d", kind);
    if (nb < 0 || (nb > 0 && !idx1)) return fail(MFT_EINVAL, "mft_add_boundary: bad index list");
    if (kind == MFT_BC_DIRICHLET && nb > 0 && !values) return fail(MFT_EINVAL, "mft_add_boundary: Dirichlet group needs a value table");
    if (kind == MFT_BC_SLIP_WALL && c->V != 4) return fail(MFT_ENOTSUP, "mft_add_boundary: slip wall is defined for Euler 2-D only");
    if (kind == MFT_BC_SLIP_WALL && nb > 0 && !normals) return fail(MFT_EINVAL, "mft_add_boundary: slip wall needs normals");
    BcGroup *g = new BcGroup();
    g->kind = kind;
    g->nb = nb;
    std::vector<int> idx(nb);
    for (int64_t j = 0; j < nb; ++j) {
        const int64_t p = idx1[j] - 1;
        if (p < 0 || p >= c->n_tot) {
            delete g;
            return fail(MFT_EINVAL, "mft_add_boundary: index %lld out of range", (long long)idx1[j]);
        }
        idx[j] = (int)p;  // caller numbering for now; remapped in finalize
    }
    int r = g->idx.upload(idx);
    if (r == 0 && normals && nb > 0) r = g->normals.upload(std::vector<double>(normals, normals + 2 * nb));
    if (r == 0 && values && nb > 0) {
        std::vector<double> aos((size_t)nb * c->V);
        for (int64_t j = 0; j < nb; ++j)
            for (int v = 0; v < c->V; ++v) aos[j * c->V + v] = values[(int64_t)v * nb + j];
        r = g->values.upload(aos);
    }
    if (r != 0) {
        delete g;
        return r;
    }
    // keep caller-numbered indices on the host for the finalize remap
    c->bcs.push_back(g);
    c->bc_caller_rows.emplace_back();
    if (kind != MFT_BC_DO_NOTHING)
        for (int64_t j = 0; j < nb; ++j) c->bc_caller_rows.back().push_back(idx1[j] - 1);
    return MFT_OK;
}

extern "C" int mft_update_boundary_values(mft_ctx *c, int group, const double *values)
{
    NEED_CTX(c);
    if (group < 0 || group >= (int)c->bcs.size()) return fail(MFT_EINVAL, "mft_update_boundary_values: group %d out of range", group);
    BcGroup *g = c->bcs[group];
    if (g->kind != MFT_BC_DIRICHLET) return fail(MFT_EINVAL, "mft_update_boundary_values: group %d is not Dirichlet", group);
    if (!values) return fail(MFT_EINVAL, "mft_update_boundary_values: values is NULL");
    std::vector<double> aos((size_t)g->nb * c->V);
    for (int64_t j = 0; j < g->nb; ++j)
        for (int v = 0; v < c->V; ++v) aos[j * c->V + v] = values[(int64_t)v * g->nb + j];
    if (g->nb > 0) CU(cudaMemcpyAsync(g->values.p, aos.data(), sizeof(double) * aos.size(), cudaMemcpyHostToDevice, c->stream));
    if (g->nb > 0 && c->bc_merged && c->bc_group_off[group] >= 0)
        CU(cudaMemcpyAsync(c->bc_values.p + c->bc_group_off[group] * c->V, aos.data(), sizeof(double) * aos.size(),
                           cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

// Dirichlet tables for the stage times of the NEXT device-resident step (mft_ssprk_step / mft_ssprk43_step): the reference's
// BC closures receive the stage time (calc_single_boundary_flux!, rbfsolver.jl:311-316: boundary_condition(..., t, ...)), and
// one step evaluates rhs! at t + dt and t + dt/2.  slot 0 = values at t + dt, slot 1 = values at t + dt/2.  The step copies the
// slot's table over the group's current table (device to device, stream-ordered, part of the captured graph) in front of the
// rhs! evaluated at that time; after the step the current table is the t + dt one, i.e. the start time of the next step.
extern "C" int mft_set_stage_boundary_values(mft_ctx *c, int group, int slot, const double *values)
{
    NEED_CTX(c);
    if (group < 0 || group >= (int)c->bcs.size()) return fail(MFT_EINVAL, "mft_set_stage_boundary_values: group %d out of range", group);
    if (slot < 0 || slot > 1) return fail(MFT_EINVAL, "mft_set_stage_boundary_values: slot must be 0 (t + dt) or 1 (t + dt/2)");
    BcGroup *g = c->bcs[group];
    if (g->kind != MFT_BC_DIRICHLET) return fail(MFT_EINVAL, "mft_set_stage_boundary_values: group %d is not Dirichlet", group);
    if (!values) return fail(MFT_EINVAL, "mft_set_stage_boundary_values: values is NULL");
    if (g->nb == 0) return MFT_OK;
    if (!g->stage_values[slot].p) {
        CHECK(g->stage_values[slot].alloc(g->nb * c->V));
        // captured steps were recorded without the table copies: start over
        for (auto &e : c->graphs)
            if (e.exec) cudaGraphExecDestroy(e.exec);
        c->graphs.clear();
    }
    std::vector<double> aos((size_t)g->nb * c->V);
    for (int64_t j = 0; j < g->nb; ++j)
        for (int v = 0; v < c->V; ++v) aos[j * c->V + v] = values[(int64_t)v * g->nb + j];
    CU(cudaMemcpyAsync(g->stage_values[slot].p, aos.data(), sizeof(double) * aos.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    g->stage_set[slot] = true;
    return MFT_OK;
}

// in front of a stage's rhs!: make the slot's Dirichlet tables the current ones (no-op unless stage tables were supplied)
static int select_stage_boundary_values(mft_ctx *c, int slot)
{
    for (size_t gi = 0; gi < c->bcs.size(); ++gi) {
        BcGroup *g = c->bcs[gi];
        if (g->kind != MFT_BC_DIRICHLET || g->nb == 0 || !g->stage_set[slot]) continue;
        const size_t bytes = sizeof(double) * (size_t)g->nb * c->V;
        CU(cudaMemcpyAsync(g->values.p, g->stage_values[slot].p, bytes, cudaMemcpyDeviceToDevice, c->stream));
        if (c->bc_merged && c->bc_group_off[gi] >= 0)
            CU(cudaMemcpyAsync(c->bc_values.p + c->bc_group_off[gi] * c->V, g->stage_values[slot].p, bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
    return MFT_OK;
}

extern "C" int mft_add_source(mft_ctx *c, int kind, const double *params, int nparams, const int64_t *colptr,
                              const int64_t *rowval, const double *nzval)
{
    NEED_CTX(c);
    if (c->finalized) return fail(MFT_EINVAL, "mft_add_source: must be called before the first compute call");
    Source *s = new Source();
    s->kind = kind;
    int r = MFT_OK;
    if (kind == MFT_SRC_HV_FLYER || kind == MFT_SRC_HV_TOMINEC) {
        if (nparams < 1 || !params) r = fail(MFT_EINVAL, "mft_add_source: hyperviscosity needs gamma");
        else {
            s->gamma = params[0];
            r = copy_csc(c, s->hv_host, colptr, rowval, nzval, "mft_add_source");
        }
    } else if (kind == MFT_SRC_UPWIND) {
        if (c->eq != MFT_EQ_EULER2D) r = fail(MFT_ENOTSUP, "mft_add_source: upwind viscosity is defined for Euler 2-D only (hyperviscosity.jl:246-247); call mft_set_equation first");
        else if (nparams < 2 || !params) r = fail(MFT_EINVAL, "mft_add_source: upwind viscosity needs c_uw, dx_avg");
        else {
            s->c_uw = params[0];
            s->dx_avg = params[1];
        }
    } else if (kind == MFT_SRC_RESIDUAL) {
        if (c->eq != MFT_EQ_EULER2D) r = fail(MFT_ENOTSUP, "mft_add_source: residual viscosity is defined for Euler 2-D only (hyperviscosity.jl:289-291); call mft_set_equation first");
        else if (nparams < 4 || !params) r = fail(MFT_EINVAL, "mft_add_source: residual viscosity needs c_rv, c_uw, dx_avg, polydeg");
        else {
            s->c_rv = params[0];
            s->c_uw = params[1];
            s->dx_avg = params[2];
            s->polydeg = (int)params[3];
            if (s->polydeg < 0 || s->polydeg > 7) r = fail(MFT_EINVAL, "mft_add_source: polydeg must be in [0,7]");
            for (auto *o : c->srcs)
                if (o->kind == MFT_SRC_RESIDUAL) r = fail(MFT_ENOTSUP, "mft_add_source: only one residual-viscosity source per ctx");
        }
    } else if (kind == MFT_SRC_IGR) {
        if (c->eq != MFT_EQ_EULER2D) r = fail(MFT_ENOTSUP, "mft_add_source: the IGR source is defined for Euler 2-D only (IGR.jl:117-119); call mft_set_equation first");
        else if (nparams < 1 || !params) r = fail(MFT_EINVAL, "mft_add_source: IGR needs alpha [, maxiter]");
        else {
            s->igr_alpha = params[0];
            s->igr_maxiter = nparams >= 2 ? (int)params[1] : 20;
            if (s->igr_maxiter < 0 || s->igr_maxiter > 1000) r = fail(MFT_EINVAL, "mft_add_source: IGR maxiter must be in [0,1000]");
        }
    } else {
        r = fail(MFT_EINVAL, "mft_add_source: unknown kind %d", kind);
    }
    if (r != 0) {
        delete s;
        return r;
    }
    c->srcs.push_back(s);
    return MFT_OK;
}

#include "mft_layout_host.inl"
#include "mft_layout_device.inl"

// per-row side table of the fused stage kernel: boundary entry (merged table) and halo routes of a row
static int build_row_aux(mft_ctx *c)
{
    const int64_t nl = c->n_local;
    std::vector<int> aux((size_t)nl + 4, -1);   // (+ padding: the stage kernel copies it in 16-byte multiples)
    std::vector<RowAux> tab;
    auto entry = [&](int row) -> RowAux & {
        if (aux[(size_t)row] < 0) {
            aux[(size_t)row] = (int)tab.size();
            tab.push_back(RowAux{-1, 0, 0});
        }
        return tab[(size_t)aux[(size_t)row]];
    };
    if (c->bc_merged)
        for (size_t j = 0; j < c->bc_idx_host.size(); ++j) {
            const int row = c->bc_idx_host[j];
            if (row >= 0 && row < nl) entry(row).bc = (int)j;
        }
    // routes sorted by row: a row that several peers hold in their halo owns a contiguous slice
    // (mft_set_halo may precede mft_finalize; the routes only exist once mft_p2p_connect has run)
    const size_t ns = c->send_peer_host.size() == c->send_rows_host.size() ? c->send_rows_host.size() : 0;
    std::vector<int> ord(ns), rpeer(ns);
    std::vector<long long> rdst(ns);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return c->send_rows_host[(size_t)a] < c->send_rows_host[(size_t)b]; });
    for (size_t q = 0; q < ns; ++q) {
        const int i = ord[q];
        const int row = c->send_rows_host[(size_t)i];
        rpeer[q] = c->send_peer_host[(size_t)i];
        rdst[q] = c->send_dst_host[(size_t)i];
        RowAux &e = entry(row);
        if (e.send == e.sbeg) e.sbeg = (int)q;
        e.send = (int)q + 1;
    }
    if (tab.empty()) tab.push_back(RowAux{-1, 0, 0});
    if (rpeer.empty()) {
        rpeer.push_back(0);
        rdst.push_back(0);
    }
    CHECK(c->row_aux.upload(aux));
    CHECK(c->row_aux_tab.upload(tab));
    CHECK(c->route_peer.upload(rpeer));
    CHECK(c->route_dst.upload(rdst));
    return MFT_OK;
}

static bool has_visc(const mft_ctx *c)
{
    for (auto *s : c->srcs)
        if (s->kind == MFT_SRC_UPWIND || s->kind == MFT_SRC_RESIDUAL) return true;
    return false;
}
extern "C" int mft_finalize(mft_ctx *c)
{
    NEED_CTX(c);
    if (c->finalized) return MFT_OK;
    if (c->eq < 0) return fail(MFT_EINVAL, "mft_finalize: mft_set_equation was not called");
    const int64_t n = c->n_tot;
    const bool trace = getenv("MFT_TRACE") != nullptr;
    auto t_phase = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {   // MFT_TRACE: wall time of the setup phases of mft_finalize
        if (!trace) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[mft] finalize: %-44s %.3f s\n", what, std::chrono::duration<double>(now - t_phase).count());
        t_phase = now;
    };
    Csr2 F, T;
    if (c->have_ell_input) {
        F = c->host_ell;
        // complete F with empty halo rows so that the transpose sees n_tot columns
        sort_rows_by_key(F, c->keys, true);
        // default summation order for the ELL input is ascending caller column index
        if (c->keys.empty()) {
            std::vector<int64_t> ident(n);
            std::iota(ident.begin(), ident.end(), (int64_t)0);
            sort_rows_by_key(F, ident, true);
        }
        transpose_rows(F, n, true, T);
        sort_rows_by_key(T, c->keys, true);
    } else {
        if (!c->host_ops[0].set || !c->host_ops[1].set) return fail(MFT_EINVAL, "mft_finalize: Dx and Dy operators were not both set");
        const HostCsc &X = c->host_ops[0], &Y = c->host_ops[1];
        if (X.colptr != Y.colptr || X.rowval != Y.rowval)
            return fail(MFT_ENOTSUP, "mft_finalize: Dx and Dy must share one sparsity pattern (compute_flux_operator builds them from the same neighbour lists)");
        csc_to_colrows(X, &Y, n, T);  // rows of D' (ascending caller row index inside each)
        transpose_rows(T, n, true, F);  // rows of D, ascending caller column index
        sort_rows_by_key(F, c->keys, true);
        sort_rows_by_key(T, c->keys, true);
    }
    lap("operator rows (transpose, summation-key sort)");
    // Within blocks of device rows, order the rows by the length of their D' row: slices of the transposed operator then
    // have near-uniform width (little padding).  Sliced-ELL kernels: blocks of 256 rows (locality is untouched inside a block,
    // measured no gain: the gathers of a warp spread).  Union-tile kernels: block = one tile, so every tile keeps exactly its
    // rows and its stencil union and only the order inside the tile changes -- simulated on the bench cloud (host replay):
    // 22.2 -> 20.7 steps per row of D' and conflict degree 1.08 -> 1.04.
    const bool tiles_on = (c->tile & 3) && c->V == 4 && c->eq == MFT_EQ_EULER2D;
    if (has_visc(c) && c->refine_order) {
        const int64_t nl = c->n_local;
        if (!c->have_perm) {
            c->perm.resize(n);
            std::iota(c->perm.begin(), c->perm.end(), 0);
            c->iperm = c->perm;
            c->have_perm = true;
        }
        auto len = [&](int32_t pt) { return T.ptr[pt + 1] - T.ptr[pt]; };
        int64_t blk = 256;
        if (tiles_on) {
            const int64_t ta = (c->tile & 1) ? (int64_t)kSlice * kTileWarps * c->tile_rows_a : (int64_t)1 << 40;
            const int64_t tb = (c->tile & 2) ? (int64_t)kSlice * kTileWarps * c->tile_rows_b : (int64_t)1 << 40;
            blk = std::min(ta, tb);   // tile sizes are 128 * {1, 2, 4}: the smaller one divides the larger
        }
        std::vector<int64_t> pos_before;   // tiles: position of a point in the order before the refinement (= along the curve)
        if (tiles_on) {
            pos_before.resize(n);
            for (int64_t d = 0; d < n; ++d) pos_before[c->perm[d]] = d;
        }
        for (int64_t b0 = 0; b0 < nl; b0 += blk) {
            const int64_t b1 = std::min(nl, b0 + blk);
            std::stable_sort(c->perm.begin() + b0, c->perm.begin() + b1, [&](int32_t x, int32_t y) { return len(x) > len(y); });
            // the length of a slice's walk depends on WHICH rows it holds, not on their order: inside every slice go back to the
            // order along the curve, so that the rows of an LDS.128 phase stay neighbours (replay: conflict degree 1.06 -> 1.05)
            if (tiles_on)
                for (int64_t s0 = b0; s0 < b1; s0 += kSlice)
                    std::sort(c->perm.begin() + s0, c->perm.begin() + std::min(b1, s0 + kSlice),
                              [&](int32_t x, int32_t y) { return pos_before[x] < pos_before[y]; });
        }
        for (int64_t d = 0; d < n; ++d) c->iperm[c->perm[d]] = (int32_t)d;
    }
    // drop halo rows of the forward operator: only owned rows are computed here
    if (c->have_perm) CHECK(c->d_perm.upload(std::vector<int>(c->perm.begin(), c->perm.end())));
    // Kernel family by stencil width (north_star: "warp-per-row or thread-per-row chosen from measured stencil width").  Measured
    // on the B200 for k = 13, 15, 20, 25, 30, 36, 42, 50 at 1M points (tools/stencil_sweep.py, profiles/r2_stencil_sweep.json):
    // union tiles 0.49 - 0.64 of the HBM peak (flux only) and 0.50 - 0.60 (full residual-viscosity rhs!), thread-per-row sliced
    // ELL 0.29 - 0.41 and 0.39 - 0.48.  The tiles win at EVERY width of the reference's range (geometry_primatives.jl:197-198:
    // k = 15 / 20 / 30 / 42), so there is no crossover to switch on: Euler 2-D always takes the tile kernels unless MFT_OPT_TILE
    // says otherwise; the only width-dependent choice left is inside the tile format (slot count <= 4095 and shared memory
    // <= 200 KB per block are checked at build / launch time and reported, not silently worked around).
    const bool tile_a = (c->tile & 1) && c->V == 4 && c->eq == MFT_EQ_EULER2D;
    const bool tile_b = (c->tile & 2) && c->V == 4 && c->eq == MFT_EQ_EULER2D;
    if (tile_a) CHECK(build_tiler(c, F, c->n_local, c->tile_rows_a, (c->tile & 4) != 0, (c->tile & 8) != 0, c->fwd_tiler));
    else CHECK(build_ell(c, F, c->n_local, true, c->fwd));
    if (!tile_a && (c->pair_rows & 2) && c->V == 4) CHECK(build_ell_pairs(c, F, c->n_local, c->fwd_pair));
    lap("forward operator layout");
    if (has_visc(c)) {
        if (tile_b) CHECK(build_tiler(c, T, c->n_local, c->tile_rows_b, (c->tile & 4) != 0, (c->tile & 8) != 0, c->tra_tiler));
        else if ((c->pair_rows & 1) && c->V == 4) CHECK(build_ell_pairs(c, T, c->n_local, c->tra_pair));
        else CHECK(build_ell(c, T, c->n_local, true, c->tra));
        lap("transposed operator layout");
        CHECK(c->g.alloc((n + 1) * 2 * c->V));  // + zero dummy record
        CU(cudaMemset(c->g.p, 0, sizeof(double) * (n + 1) * 2 * c->V));
    }
    for (auto *s : c->srcs) {
        if (s->kind == MFT_SRC_HV_FLYER || s->kind == MFT_SRC_HV_TOMINEC) {
            Csr2 HT, H;
            csc_to_colrows(s->hv_host, nullptr, n, HT);
            transpose_rows(HT, n, false, H);
            sort_rows_by_key(H, c->keys, false);
            CHECK(build_ell(c, H, c->n_local, false, s->hv));
            s->hv_host = HostCsc();
        }
        if (s->kind == MFT_SRC_IGR) {
            using mft_igr::IgrArgs;
            using mft_igr::IgrScalars;
            CHECK(build_ell(c, F, c->n_local, true, s->igr_op));
            const int64_t nl = c->n_local, np1 = n + 1;
            CHECK(s->igr_vec.alloc(4 * nl + 4 * np1));
            CU(cudaMemset(s->igr_vec.p, 0, sizeof(double) * (4 * nl + 4 * np1)));
            const int nblocks = grid_for(nl, mft_igr::kBlock);
            CHECK(s->igr_partial.alloc(nblocks));
            CHECK(s->igr_ticket.alloc(1));
            CU(cudaMemset(s->igr_ticket.p, 0, sizeof(unsigned int)));
            IgrScalars init{};
            init.prev_res = 1.0;
            init.maxiter = s->igr_maxiter;
            init.done = 1;
            CHECK(s->igr_scalars.upload(std::vector<IgrScalars>(1, init)));
            IgrArgs &A = s->igr_args;
            A.blob = s->igr_op.blob.p;
            A.off = s->igr_op.off.p;
            A.n_rows = nl;
            A.u = c->u.p;
            A.du = c->du.p;
            A.alpha = s->igr_alpha;
            double *v = s->igr_vec.p;
            A.rho_inv = v;
            A.b = v + nl;
            A.r = v + 2 * nl;
            A.c = v + 3 * nl;
            A.x = v + 4 * nl;
            A.p = v + 4 * nl + np1;
            A.t = v + 4 * nl + 2 * np1;
            A.partial = s->igr_partial.p;
            A.ticket = s->igr_ticket.p;
            A.S = s->igr_scalars.p;
        }
        if (s->kind == MFT_SRC_RESIDUAL) {
            c->nslots = s->polydeg + 1;
            c->hist.resize(c->nslots);
            for (auto &h : c->hist) {
                CHECK(h.alloc(n * c->V));
                CU(cudaMemset(h.p, 0, sizeof(double) * n * c->V));
            }
            c->time_history.assign(c->nslots, 0.0);
            c->time_weights.assign(c->nslots, 0.0);
            CHECK(c->approx_du.alloc(n * c->V));
            CU(cudaMemset(c->approx_du.p, 0, sizeof(double) * n * c->V));
        }
    }
    if (c->diagnostics && has_visc(c)) {
        CHECK(c->eps.alloc(n));
        CHECK(c->eps_uw.alloc(n));
        CHECK(c->eps_rv.alloc(n));
        CHECK(c->eps_c.alloc(n));
        CHECK(c->residual.alloc(n * c->V));
        CU(cudaMemset(c->eps.p, 0, sizeof(double) * n));
        CU(cudaMemset(c->eps_uw.p, 0, sizeof(double) * n));
        CU(cudaMemset(c->eps_rv.p, 0, sizeof(double) * n));
        CU(cudaMemset(c->eps_c.p, 0, sizeof(double) * n));
        CU(cudaMemset(c->residual.p, 0, sizeof(double) * n * c->V));
    }
    // boundary indices: caller numbering -> device rows
    if (c->have_perm) {
        for (auto *g : c->bcs) {
            if (g->nb == 0) continue;
            std::vector<int> idx(g->nb);
            CU(cudaMemcpy(idx.data(), g->idx.p, sizeof(int) * g->nb, cudaMemcpyDeviceToHost));
            for (auto &i : idx) i = c->iperm[i];
            CU(cudaMemcpy(g->idx.p, idx.data(), sizeof(int) * g->nb, cudaMemcpyHostToDevice));
        }
    }
    // one merged table for all groups when no point is in two groups (then the group order cannot matter)
    {
        std::vector<int> kind, idx;
        std::vector<double> nrm, val;
        c->bc_group_off.assign(c->bcs.size(), -1);
        for (size_t gi = 0; gi < c->bcs.size(); ++gi) {
            BcGroup *g = c->bcs[gi];
            if (g->nb == 0 || g->kind == MFT_BC_DO_NOTHING) continue;
            c->bc_group_off[gi] = (int64_t)idx.size();
            std::vector<int> gidx(g->nb);
            CU(cudaMemcpy(gidx.data(), g->idx.p, sizeof(int) * g->nb, cudaMemcpyDeviceToHost));
            std::vector<double> gn(2 * g->nb, 0.0), gv((size_t)g->nb * c->V, 0.0);
            if (g->normals.p) CU(cudaMemcpy(gn.data(), g->normals.p, sizeof(double) * 2 * g->nb, cudaMemcpyDeviceToHost));
            if (g->values.p) CU(cudaMemcpy(gv.data(), g->values.p, sizeof(double) * g->nb * c->V, cudaMemcpyDeviceToHost));
            idx.insert(idx.end(), gidx.begin(), gidx.end());
            kind.insert(kind.end(), (size_t)g->nb, g->kind);
            nrm.insert(nrm.end(), gn.begin(), gn.end());
            val.insert(val.end(), gv.begin(), gv.end());
        }
        std::vector<int> sorted(idx);
        std::sort(sorted.begin(), sorted.end());
        c->bc_merged = std::adjacent_find(sorted.begin(), sorted.end()) == sorted.end();
        if (c->bc_merged) {
            c->bc_total = (int64_t)idx.size();
            c->bc_idx_host = idx;
            CHECK(c->bc_kind.upload(kind));
            CHECK(c->bc_idx.upload(idx));
            CHECK(c->bc_normals.upload(nrm));
            CHECK(c->bc_values.upload(val));
        }
    }
    {
        std::vector<int64_t> rows;
        for (auto &v : c->bc_caller_rows) rows.insert(rows.end(), v.begin(), v.end());
        for (int64_t h = c->n_local; h < n; ++h) rows.push_back(h);
        std::sort(rows.begin(), rows.end());
        rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        c->touched_caller = rows;
        std::vector<int> dev(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) dev[i] = c->have_perm ? c->iperm[rows[i]] : (int)rows[i];
        CHECK(c->touched_dev.upload(dev));
        CHECK(c->touched_buf.alloc(std::max<int64_t>(1, (int64_t)rows.size()) * c->V));
        CU(cudaHostAlloc((void **)&c->touched_host, sizeof(double) * std::max<size_t>(1, rows.size()) * c->V, cudaHostAllocDefault));
    }
    CHECK(c->uprev.alloc(n * c->V));
    CHECK(build_row_aux(c));
    // free host staging
    c->host_ops[0] = HostCsc();
    c->host_ops[1] = HostCsc();
    c->host_ell = Csr2();
    c->finalized = true;
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// state transfer
// ------------------------------------------------------------------------------------------------------
static int upload_soa(mft_ctx *c, const double *const *soa, double *dst_aos)
{
    const int64_t n = c->n_tot;
    // Multi-rank: only the owned rows are taken from the caller.  The halo tail belongs to the exchange -- with
    // peer-memory puts a neighbour may already have written the next halo block while this upload is still queued.
    const int64_t n_up = c->nranks > 1 ? c->n_local : n;
    for (int v = 0; v < c->V; ++v) {
        if (!soa || !soa[v]) return fail(MFT_EINVAL, "state component %d is NULL", v);
        CU(cudaMemcpyAsync(c->stage_soa.p + (int64_t)v * n, soa[v], sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    }
    ScopedTimer t(c, MFT_K_OTHER);
    if (c->V == 4)
        k_pack<4><<<grid_for(n_up, 256), 256, 0, c->stream>>>(c->stage_soa.p, n, c->d_perm.p, reinterpret_cast<Vec<4> *>(dst_aos), n_up);
    else
        k_pack<1><<<grid_for(n_up, 256), 256, 0, c->stream>>>(c->stage_soa.p, n, c->d_perm.p, reinterpret_cast<Vec<1> *>(dst_aos), n_up);
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

static int download_soa(mft_ctx *c, const double *src_aos, double *const *soa)
{
    const int64_t n = c->n_tot;
    {
        ScopedTimer t(c, MFT_K_OTHER);
        if (c->V == 4)
            k_unpack<4><<<grid_for(n, 256), 256, 0, c->stream>>>(reinterpret_cast<const Vec<4> *>(src_aos), c->d_perm.p, c->stage_soa.p, n, n);
        else
            k_unpack<1><<<grid_for(n, 256), 256, 0, c->stream>>>(reinterpret_cast<const Vec<1> *>(src_aos), c->d_perm.p, c->stage_soa.p, n, n);
        c->launches++;
        LAUNCH_CHECK();
    }
    for (int v = 0; v < c->V; ++v) {
        if (!soa || !soa[v]) return fail(MFT_EINVAL, "state component %d is NULL", v);
        CU(cudaMemcpyAsync(soa[v], c->stage_soa.p + (int64_t)v * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    return MFT_OK;
}

// fused step: rows that exceeded the one-pass norms since the last check (loud, never silent: MFT_ENORMS)
static int check_norm_misses(mft_ctx *c)
{
    if (!c->fused_used || !c->norm_miss.p) return MFT_OK;
    unsigned long long m = 0;
    CU(cudaMemcpy(&m, c->norm_miss.p, sizeof m, cudaMemcpyDeviceToHost));
    if (m > c->norm_miss_seen) {
        const unsigned long long d = m - c->norm_miss_seen;
        c->norm_miss_seen = m;
        return fail(MFT_ENORMS, "fused step: %llu row(s) exceeded the one-pass ode_maximum statistic (a rounding tie decided the lexicographic "
                                "order); the affected stages used norms that differ from the reference's -- re-run with MFT_OPT_FUSED_STEP = 0",
                    d);
    }
    return MFT_OK;
}

extern "C" int mft_upload_state(mft_ctx *c, const double *const *u_soa)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    CHECK(upload_soa(c, u_soa, c->u.p));
    CU(cudaStreamSynchronize(c->stream));
    c->have_fsal = false;
    return MFT_OK;
}
extern "C" int mft_download_state(mft_ctx *c, double *const *u_soa)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    CHECK(download_soa(c, c->u.p, u_soa));
    CU(cudaStreamSynchronize(c->stream));
    return check_norm_misses(c);
}
extern "C" int mft_download_du(mft_ctx *c, double *const *du_soa)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    CHECK(download_soa(c, c->du.p, du_soa));
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// halo exchange (NCCL send/recv straight from/to device buffers; receive lands in the halo tail)
// ------------------------------------------------------------------------------------------------------
// peer-memory exchange, split so that independent work can be queued between the put and the wait
template <int W>
static int p2p_put(mft_ctx *c, double *field)
{
    ScopedTimer t(c, MFT_K_OTHER);
    P2PLocal *L = reinterpret_cast<P2PLocal *>(c->p2p_local.p);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(128, (c->n_send + 255) / 256));
    k_p2p_put<W><<<grid, 256, 0, c->stream>>>(c->peers_dev, L, W == 8 ? 1 : 0, reinterpret_cast<const Vec<W> *>(field),
                                             c->send_rows.p, c->send_peer.p, c->send_dst.p, c->n_send);
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}
static int p2p_wait(mft_ctx *c, int F)
{
    ScopedTimer t(c, MFT_K_OTHER);
    k_p2p_wait<<<1, 32, 0, c->stream>>>(c->peers_dev, reinterpret_cast<P2PLocal *>(c->p2p_local.p), F);
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

template <int W>
static int halo_exchange(mft_ctx *c, double *field /* AoS, W doubles per point */)
{
    if (c->nranks <= 1 || (c->n_send == 0 && c->n_halo == 0)) return MFT_OK;
    if (c->p2p) {
        CHECK(p2p_put<W>(c, field));
        return p2p_wait(c, W == 8 ? 1 : 0);
    }
    if (!c->comm) return fail(MFT_EINVAL, "halo exchange requested but mft_comm_init was not called");
    if (c->n_send > 0) {
        ScopedTimer t(c, MFT_K_OTHER);
        k_halo_pack<W><<<grid_for(c->n_send, 256), 256, 0, c->stream>>>(reinterpret_cast<const Vec<W> *>(field), c->send_rows.p,
                                                                     reinterpret_cast<Vec<W> *>(c->send_buf.p), c->n_send);
        c->launches++;
        LAUNCH_CHECK();
    }
    NcclApi *N = c->nccl;
    if (N->groupStart() != 0) return fail(MFT_ENCCL, "ncclGroupStart failed");
    for (size_t p = 0; p < c->peers.size(); ++p) {
        const int64_t ns = c->send_off[p + 1] - c->send_off[p];
        const int64_t nr = c->recv_off[p + 1] - c->recv_off[p];
        if (ns > 0 && N->send(c->send_buf.p + c->send_off[p] * W, (size_t)ns * W, NCCL_DOUBLE, c->peers[p], c->comm, c->stream) != 0)
            return fail(MFT_ENCCL, "ncclSend failed: %s", N->lastError(c->comm));
        if (nr > 0 && N->recv(field + (c->n_local + c->recv_off[p]) * W, (size_t)nr * W, NCCL_DOUBLE, c->peers[p], c->comm, c->stream) != 0)
            return fail(MFT_ENCCL, "ncclRecv failed: %s", N->lastError(c->comm));
    }
    if (N->groupEnd() != 0) return fail(MFT_ENCCL, "ncclGroupEnd failed: %s", N->lastError(c->comm));
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// kernels launch helpers
// ------------------------------------------------------------------------------------------------------
static int launch_boundary(mft_ctx *c, bool write_du)
{
    if (c->bc_merged) {
        if (c->bc_total == 0) return MFT_OK;
        ScopedTimer t(c, MFT_K_BC);
        BcMergedArgs a{c->bc_total, c->bc_kind.p, c->bc_idx.p, c->bc_normals.p, c->bc_values.p, c->u.p, write_du ? c->du.p : nullptr};
        if (c->V == 4)
            k_boundary_merged<4><<<grid_for(a.nb, 128), 128, 0, c->stream>>>(a);
        else
            k_boundary_merged<1><<<grid_for(a.nb, 128), 128, 0, c->stream>>>(a);
        c->launches++;
        LAUNCH_CHECK();
        return MFT_OK;
    }
    for (auto *g : c->bcs) {
        if (g->nb == 0 || g->kind == MFT_BC_DO_NOTHING) continue;
        ScopedTimer t(c, MFT_K_BC);
        BcArgs a{g->kind, g->nb, g->idx.p, g->normals.p, g->values.p, c->u.p, write_du ? c->du.p : nullptr};
        if (c->V == 4)
            k_boundary<4><<<grid_for(g->nb, 128), 128, 0, c->stream>>>(a);
        else
            k_boundary<1><<<grid_for(g->nb, 128), 128, 0, c->stream>>>(a);
        c->launches++;
        LAUNCH_CHECK();
    }
    return MFT_OK;
}

// dynamic shared memory above 48 KB needs an opt-in per kernel
template <typename K>
static int ensure_smem(mft_ctx *c, K kernel, int bytes)
{
    const void *key = reinterpret_cast<const void *>(kernel);
    if (std::find(c->smem_configured.begin(), c->smem_configured.end(), key) != c->smem_configured.end()) return MFT_OK;
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    c->smem_configured.push_back(key);
    (void)bytes;
    return MFT_OK;
}

static inline int warp_buf_bytes(const DevEll &e, bool stage_w)
{
    const int per_col = stage_w ? e.colb : kSlice * 4;
    return ((std::max(e.maxw, 1) * per_col + 127) / 128) * 128;
}

// Whole-slice staging keeps 4 CTAs/SM only while a warp's blob is <= ~14 KB (k <= 22); wider stencils stage the index
// block only (measured: stencil sweep in profiles/), which keeps occupancy and the L1 carve-out.
static inline bool stage_whole_slice(const mft_ctx *c, const DevEll &e, bool want)
{
    return want && e.maxw * e.colb <= 14 * 1024;
}

template <int V, int EQ, bool EX, bool DF, int VI>
static int launch_pass_a_k(mft_ctx *c, const PassAArgs &a, int grid, int smem)
{
    const bool stage_w = stage_whole_slice(c, c->fwd, c->stage_w);
    if constexpr (V == 4 && EX && DF) {
        // uniform 20-wide forward operator (degree-3 default stencil): single-sweep exact kernel
        if (c->kfix_ok && c->fwd.maxw == 20 && c->fwd.ncols_total == (int64_t)20 * c->fwd.nslices && c->stage_w) {
            CHECK(ensure_smem(c, k_pass_a<V, EQ, EX, DF, VI, true, 20>, smem));
            k_pass_a<V, EQ, EX, DF, VI, true, 20><<<grid, 128, smem, c->stream>>>(a);
            return MFT_OK;
        }
    }
    if (stage_w) {
        CHECK(ensure_smem(c, k_pass_a<V, EQ, EX, DF, VI, true>, smem));
        k_pass_a<V, EQ, EX, DF, VI, true><<<grid, 128, smem, c->stream>>>(a);
    } else {
        CHECK(ensure_smem(c, k_pass_a<V, EQ, EX, DF, VI, false>, smem));
        k_pass_a<V, EQ, EX, DF, VI, false><<<grid, 128, smem, c->stream>>>(a);
    }
    return MFT_OK;
}

template <int V, int EQ>
static int launch_pass_a_t(mft_ctx *c, const PassAArgs &a, bool do_flux, int visc)
{
    const int grid = (int)((a.n_slices + 3) / 4);
    const int smem = 4 * a.buf_bytes;
#define PA(EX, DF, VI) CHECK((launch_pass_a_k<V, EQ, EX, DF, VI>(c, a, grid, smem)))
    if constexpr (V == 4) {
        if (c->exact) {
            if (do_flux && visc == VISC_NONE) PA(true, true, VISC_NONE);
            else if (do_flux && visc == VISC_UPWIND) PA(true, true, VISC_UPWIND);
            else if (do_flux && visc == VISC_RESIDUAL) PA(true, true, VISC_RESIDUAL);
            else if (!do_flux && visc == VISC_UPWIND) PA(true, false, VISC_UPWIND);
            else if (!do_flux && visc == VISC_RESIDUAL) PA(true, false, VISC_RESIDUAL);
            else return fail(MFT_EINVAL, "pass A: nothing to do");
        } else {
            if (do_flux && visc == VISC_NONE) PA(false, true, VISC_NONE);
            else if (do_flux && visc == VISC_UPWIND) PA(false, true, VISC_UPWIND);
            else if (do_flux && visc == VISC_RESIDUAL) PA(false, true, VISC_RESIDUAL);
            else if (!do_flux && visc == VISC_UPWIND) PA(false, false, VISC_UPWIND);
            else if (!do_flux && visc == VISC_RESIDUAL) PA(false, false, VISC_RESIDUAL);
            else return fail(MFT_EINVAL, "pass A: nothing to do");
        }
    } else {
        if (!do_flux || visc != VISC_NONE) return fail(MFT_ENOTSUP, "viscosity sources need Euler 2-D");
        if (c->exact) PA(true, true, VISC_NONE);
        else PA(false, true, VISC_NONE);
    }
#undef PA
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

// per-warp staging buffer: the step words, plus (staged weights) the R compact weight blocks of one direction
static TileROp tiler_view(const DevTileR &e, bool stage_w, int pf_slices = 0)
{
    const int bytes = std::max(e.maxW, 1) * kSlice * 2 + (stage_w ? e.R * std::max(e.maxL, 1) * kSlice * 8 : 0);
    return TileROp{e.blob.p, e.boff.p, e.wl.p, e.uoff.p, e.ulist.p, e.uslot.p, e.sstride, e.ncopy, ((bytes + 127) / 128) * 128, pf_slices / kTileWarps,
                   nullptr, 0};
}
// fused step on several GPUs: interior tiles first, band tiles wait for the halo inside the kernel
static void tiler_band_order(const mft_ctx *c, const DevTileR &e, TileROp &t)
{
    if (c->fused_active && c->p2p && c->nranks > 1 && e.order.p) {
        t.order = e.order.p;
        t.n_free = e.n_free;
    }
}

// staged weights keep at least `min_blocks` blocks per SM resident; otherwise stream them
static bool tiler_stage(const mft_ctx *c, const DevTileR &e, int narrays, bool want, int min_blocks)
{
    // measured (profiles/README.md): with the L2 prefetch of later tiles, streamed weights beat staged ones (pass A 152 vs
    // 156 us), so the tile kernels stage only on explicit request (MFT_OPT_STAGE_WEIGHTS bit 3)
    if (!want || !c->exact || !c->stage_force) return false;
    min_blocks = 1;
    const TileROp t = tiler_view(e, true);
    const int smem = narrays * e.ncopy * t.sstride * 16 + kTileWarps * t.buf_bytes + 1024;
    return smem * min_blocks <= 224 * 1024;
}

template <int R>
static int launch_pass_a_tiler(mft_ctx *c, const PassAArgs &a0, bool do_flux, int visc)
{
    PassAArgs a = a0;
    const DevTileR &e = c->fwd_tiler;
    const bool stage = tiler_stage(c, e, 3, c->stage_w, R == 1 ? 4 : R == 2 ? 3 : 2);
    TileROp t = tiler_view(e, stage, c->pf_dist);
    tiler_band_order(c, e, t);
    a.n_slices = e.nslices;
    const int grid = e.ntiles;
    const int smem = 3 * e.ncopy * t.sstride * 16 + kTileWarps * t.buf_bytes;
    const bool use_pdl = c->fused_active && c->pdl && c->pdl_next;
    if (smem > 200 * 1024) return fail(MFT_ENOTSUP, "union tile: %d bytes of shared memory per block", smem);
#define PAR(EX, DF, VI, ST)                                                                         \
    do {                                                                                            \
        CHECK(ensure_smem(c, k_pass_a_tiler<R, EX, DF, VI, ST>, smem));                             \
        CU(launch_k(k_pass_a_tiler<R, EX, DF, VI, ST>, grid, kTileWarps * 32, smem, c->stream, use_pdl, a, t)); \
    } while (0)
#define PAR2(DF, VI)                                        \
    do {                                                    \
        if (!c->exact) PAR(false, DF, VI, false);           \
        else if (stage) PAR(true, DF, VI, true);            \
        else PAR(true, DF, VI, false);                      \
    } while (0)
    if (do_flux && visc == VISC_NONE) PAR2(true, VISC_NONE);
    else if (do_flux && visc == VISC_UPWIND) PAR2(true, VISC_UPWIND);
    else if (do_flux && visc == VISC_RESIDUAL) PAR2(true, VISC_RESIDUAL);
    else if (!do_flux && visc == VISC_UPWIND) PAR2(false, VISC_UPWIND);
    else if (!do_flux && visc == VISC_RESIDUAL) PAR2(false, VISC_RESIDUAL);
    else return fail(MFT_EINVAL, "pass A: nothing to do");
#undef PAR2
#undef PAR
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

template <int R>
static int launch_pass_b_tiler(mft_ctx *c)
{
    const DevTileR &e = c->tra_tiler;
    const bool stage = tiler_stage(c, e, 4, c->stage_w_b, R == 1 ? 4 : R == 2 ? 3 : 2);
    TileROp t = tiler_view(e, stage, c->pf_dist);
    tiler_band_order(c, e, t);
    PassBTileArgs a{c->g.p, c->du.p, c->n_local, e.nslices, t.order ? c->peers_dev_buf.p : nullptr,
                    reinterpret_cast<P2PLocal *>(c->p2p_local.p)};
    const int smem = 4 * e.ncopy * t.sstride * 16 + kTileWarps * t.buf_bytes;
    if (smem > 200 * 1024) return fail(MFT_ENOTSUP, "union tile: %d bytes of shared memory per block", smem);
    const bool use_pdl = c->fused_active && c->pdl && c->pdl_next;
#define PBR(EX, ST)                                                                                  \
    do {                                                                                             \
        CHECK(ensure_smem(c, k_pass_b_tiler<R, EX, ST>, smem));                                      \
        CU(launch_k(k_pass_b_tiler<R, EX, ST>, e.ntiles, kTileWarps * 32, smem, c->stream, use_pdl, a, t)); \
    } while (0)
    if (!c->exact) PBR(false, false);
    else if (stage) PBR(true, true);
    else PBR(true, false);
#undef PBR
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

static int launch_pass_a(mft_ctx *c, bool do_flux, int visc, const Source *s, bool accumulate)
{
    ScopedTimer t(c, MFT_K_PASS_A);
    PassAArgs a{};
    const bool use_tiler = c->fwd_tiler.ready();
    const bool use_tile = use_tiler;
    const bool use_pair = !use_tile && do_flux && c->V == 4 && c->fwd_pair.blob.p != nullptr;
    if (!use_tile) {
        a.op = c->fwd.view();
        a.n_slices = c->fwd.nslices;
        a.buf_bytes = warp_buf_bytes(c->fwd, stage_whole_slice(c, c->fwd, c->stage_w));
        a.two_phase = c->two_phase && c->exact && stage_whole_slice(c, c->fwd, c->stage_w);
        if (a.two_phase) a.buf_bytes = ((std::max(c->fwd.maxw, 1) * kSlice * 12 + 127) / 128) * 128;
    }
    a.pf_dist = c->pf_dist;
    a.dummy = (int)c->n_tot;
    a.u = c->u.p;
    a.du = c->du.p;
    a.g = c->g.p;
    a.approx_du = c->approx_du.p;
    a.norms = c->stats.p + 2 * c->V;
    a.norm_parts = 0;
    a.norm_lex = c->max_lex;
    a.norms_out = nullptr;
    if (c->nranks > 1 && visc == VISC_RESIDUAL) {
        a.norms = c->gather_buf.p + (int64_t)c->nranks * c->V;
        a.norm_parts = c->nranks;
        a.norms_out = c->stats.p + 2 * c->V;
    }
    a.n_rows = c->n_local;
    a.eqp0 = c->eqp[0];
    a.eqp1 = c->eqp[1];
    a.c_uw = s ? s->c_uw : 0.0;
    a.c_rv = s ? s->c_rv : 0.0;
    a.dx_avg = s ? s->dx_avg : 0.0;
    a.success_iter_zero = c->success_iter == 0;
    a.accumulate = accumulate ? 1 : 0;
    if (c->fused_active && use_tiler) {
        // fused step: the stage kernel left the norms (one GPU) or the ranks' records (several GPUs: block 0 merges them)
        if (visc == VISC_RESIDUAL) {
            a.norms = c->stats.p + 2 * c->V;
            a.norm_parts = 0;
            a.norms_out = nullptr;
            a.stats = c->stats.p;
            a.norm_miss = c->norm_miss.p;
        }
        if (c->p2p && c->nranks > 1) {
            a.P = c->peers_dev_buf.p;
            a.L = reinterpret_cast<P2PLocal *>(c->p2p_local.p);
            a.aux = c->row_aux.p;
            a.rows = c->row_aux_tab.p;
            a.route_peer = c->route_peer.p;
            a.route_dst = c->route_dst.p;
            const double ng = (double)c->n_global;
            a.divisor = c->mean_div_vn ? (double)c->V * ng : ng;
            a.norm_merge = visc == VISC_RESIDUAL ? 1 : 0;
        }
    }
    if (c->diagnostics && visc != VISC_NONE) {
        a.eps = c->eps.p;
        a.eps_uw = c->eps_uw.p;
        a.eps_rv = c->eps_rv.p;
        a.eps_c = c->eps_c.p;
        a.residual = c->residual.p;
    }
    if (use_tiler) {
        const int R = c->fwd_tiler.R;
        return R == 1 ? launch_pass_a_tiler<1>(c, a, do_flux, visc) : R == 2 ? launch_pass_a_tiler<2>(c, a, do_flux, visc) : launch_pass_a_tiler<4>(c, a, do_flux, visc);
    }
    if (use_pair) {
        const DevEll &e = c->fwd_pair;
        a.op = e.view();
        a.n_slices = e.nslices;
        a.buf_bytes = ((std::max(e.maxw, 1) * kSlice * 4 + 127) / 128) * 128;
        const int grid = (int)((a.n_slices + 3) / 4);
        const int smem = 4 * a.buf_bytes;
#define PAP(EX, VI)                                                                      \
    do {                                                                                 \
        CHECK(ensure_smem(c, k_pass_a_pair<4, EQ_EULER2D, EX, VI>, smem));               \
        k_pass_a_pair<4, EQ_EULER2D, EX, VI><<<grid, 128, smem, c->stream>>>(a);         \
    } while (0)
        if (c->exact) {
            if (visc == VISC_NONE) PAP(true, VISC_NONE);
            else if (visc == VISC_UPWIND) PAP(true, VISC_UPWIND);
            else PAP(true, VISC_RESIDUAL);
        } else {
            if (visc == VISC_NONE) PAP(false, VISC_NONE);
            else if (visc == VISC_UPWIND) PAP(false, VISC_UPWIND);
            else PAP(false, VISC_RESIDUAL);
        }
#undef PAP
        c->launches++;
        LAUNCH_CHECK();
        return MFT_OK;
    }
    if (c->V == 4) return launch_pass_a_t<4, EQ_EULER2D>(c, a, do_flux, visc);
    return launch_pass_a_t<1, EQ_ADVECTION2D>(c, a, do_flux, visc);
}

static int launch_pass_b(mft_ctx *c)
{
    ScopedTimer t(c, MFT_K_PASS_B);
    if (c->tra_tiler.ready()) {
        const int R = c->tra_tiler.R;
        return R == 1 ? launch_pass_b_tiler<1>(c) : R == 2 ? launch_pass_b_tiler<2>(c) : launch_pass_b_tiler<4>(c);
    }
    if (c->tra_pair.blob.p) {
        const DevEll &e = c->tra_pair;
        PassBPairArgs a{e.view(), c->g.p, c->du.p, c->n_local, e.nslices, ((std::max(e.maxw, 1) * kSlice * 4 + 127) / 128) * 128, (int)c->n_tot};
        const int grid = (int)((a.n_slices + 3) / 4);
        const int smem = 4 * a.buf_bytes;
        if (c->exact) {
            CHECK(ensure_smem(c, k_pass_b_pair<4, true>, smem));
            k_pass_b_pair<4, true><<<grid, 128, smem, c->stream>>>(a);
        } else {
            CHECK(ensure_smem(c, k_pass_b_pair<4, false>, smem));
            k_pass_b_pair<4, false><<<grid, 128, smem, c->stream>>>(a);
        }
        c->launches++;
        LAUNCH_CHECK();
        return MFT_OK;
    }
    const bool stage_b = stage_whole_slice(c, c->tra, c->stage_w_b);
    PassBArgs a{c->tra.view(), c->g.p, c->du.p, c->n_local, c->tra.nslices, warp_buf_bytes(c->tra, stage_b), c->pf_dist, (int)c->n_tot};
    const int grid = (int)((a.n_slices + 3) / 4);
    const int smem = 4 * a.buf_bytes;
#define PB(EX, ST)                                                   \
    do {                                                             \
        CHECK(ensure_smem(c, k_pass_b<4, EX, ST>, smem));            \
        k_pass_b<4, EX, ST><<<grid, 128, smem, c->stream>>>(a);      \
    } while (0)
    if (c->exact) {
        if (stage_b) PB(true, true); else PB(true, false);
    } else {
        if (stage_b) PB(false, true); else PB(false, false);
    }
#undef PB
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

static int launch_spmv(mft_ctx *c, const Source *s)
{
    ScopedTimer t(c, MFT_K_OTHER);
    const bool stage_s = stage_whole_slice(c, s->hv, c->stage_w);
    SpmvArgs a{s->hv.view(), c->u.p, c->du.p, c->n_local, s->hv.nslices, warp_buf_bytes(s->hv, stage_s), c->pf_dist, (int)c->n_tot, -s->gamma};
    const int grid = (int)((a.n_slices + 3) / 4);
    const int smem = 4 * a.buf_bytes;
    if (smem > 200 * 1024) return fail(MFT_ENOTSUP, "hyperviscosity operator rows too long (%d) for the shared-memory staging", s->hv.maxw);
#define SP(VV, EX, ST)                                                 \
    do {                                                               \
        CHECK(ensure_smem(c, k_spmv_accum<VV, EX, ST>, smem));         \
        k_spmv_accum<VV, EX, ST><<<grid, 128, smem, c->stream>>>(a);   \
    } while (0)
    if (c->V == 4) {
        if (c->exact) { if (stage_s) SP(4, true, true); else SP(4, true, false); }
        else { if (stage_s) SP(4, false, true); else SP(4, false, false); }
    } else {
        if (c->exact) { if (stage_s) SP(1, true, true); else SP(1, true, false); }
        else { if (stage_s) SP(1, false, true); else SP(1, false, false); }
    }
#undef SP
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

// ode_mean + ode_maximum of |u - mean| on the post-BC state (hyperviscosity.jl:305-311): two launches,
// each finishing its own cross-block reduction (last-block ticket)
static int launch_norms(mft_ctx *c)
{
    ScopedTimer t(c, MFT_K_REDUCE);
    const int V = 4;
    const int64_t n = c->n_local;
    const Vec<4> *u = reinterpret_cast<const Vec<4> *>(c->u.p);
    double *sum = c->stats.p, *mean = c->stats.p + V, *norms = c->stats.p + 2 * V;
    const double divisor = c->mean_div_vn ? (double)V * (double)n : (double)n;
    k_sum_mean<4><<<c->red_blocks, 256, 0, c->stream>>>(u, n, c->partial.p, c->ticket.p, divisor, sum);
    if (c->max_lex)
        k_maxdev_norms<4, true><<<c->red_blocks, 256, 0, c->stream>>>(u, n, sum, 1, divisor, c->partial.p, c->ticket.p + 1, norms, 1, mean);
    else
        k_maxdev_norms<4, false><<<c->red_blocks, 256, 0, c->stream>>>(u, n, sum, 1, divisor, c->partial.p, c->ticket.p + 1, norms, 1, mean);
    c->launches += 2;
    LAUNCH_CHECK();
    return MFT_OK;
}

static int launch_norms_multi(mft_ctx *c);
static int p2p_norms_part(mft_ctx *c, int part);

// SourceIGR functor call (IGR.jl:211-239): right-hand side + CG start, maxiter CG iterations of four launches each
// (device-side control: converged iterations return at once), then the sigma-flux accumulation.  Fixed launch sequence.
static int launch_igr(mft_ctx *c, Source *s)
{
    using namespace mft_igr;
    if (c->nranks > 1) return fail(MFT_ENOTSUP, "the IGR source is single-GPU (its linear solve needs halo refreshes and global dot products per iteration)");
    ScopedTimer tm(c, MFT_K_OTHER);
    const IgrArgs &A = s->igr_args;
    const int grid = grid_for(A.n_rows, kBlock);
    k_igr_rhs<<<grid, kBlock, 0, c->stream>>>(A);
    for (int it = 0; it < s->igr_maxiter; ++it) {
        k_igr_dir<<<grid, kBlock, 0, c->stream>>>(A);
        k_igr_grad<<<grid, kBlock, 0, c->stream>>>(A);
        k_igr_apply<<<grid, kBlock, 0, c->stream>>>(A);
        k_igr_update<<<grid, kBlock, 0, c->stream>>>(A);
    }
    k_igr_flux<<<grid, kBlock, 0, c->stream>>>(A);
    c->launches += 2 + 4 * s->igr_maxiter;
    LAUNCH_CHECK();
    return MFT_OK;
}

// one source functor call on the resident state
static int apply_source_dev(mft_ctx *c, Source *s)
{
    static const char *const kLabels[] = {"calc SourceHyperviscosityFlyer", "calc SourceHyperviscosityTominec",
                                          "calc SourceUpwindViscosityTominec", "calc SourceResidualViscosityTominec", "calc SourceIGR"};
    NvtxRange r(s->kind >= 0 && s->kind <= MFT_SRC_IGR ? kLabels[s->kind] : "calc source");
    if (s->kind == MFT_SRC_IGR) return launch_igr(c, s);
    if (s->kind == MFT_SRC_HV_FLYER || s->kind == MFT_SRC_HV_TOMINEC) return launch_spmv(c, s);
    const int visc = s->kind == MFT_SRC_UPWIND ? VISC_UPWIND : VISC_RESIDUAL;
    if (visc == VISC_RESIDUAL) CHECK(c->nranks > 1 ? launch_norms_multi(c) : launch_norms(c));
    CHECK(launch_pass_a(c, false, visc, s, false));
    CHECK(halo_exchange<8>(c, c->g.p));
    return launch_pass_b(c);
}

// Trixi.rhs! on the resident state: du <- rhs(u), u gets the strong BCs
static int rhs_device(mft_ctx *c, double t)
{
    (void)t;  // Dirichlet tables are refreshed by the caller (mft_update_boundary_values) when time-dependent
    // update_halos! (parallel_rbfsolver.jl:98-101) happens before the BC pass in the reference; BC points are owned
    // points, and halo copies of boundary points must carry the BC-imposed value the owner computes, so the
    // exchange runs after the owner applied its BCs.
    NvtxRange rhs_range("rhs!");
    {
        NvtxRange r("boundary flux");
        CHECK(launch_boundary(c, false));  // pass 1: du is formed from 0 below, only u needs writing
    }
    const bool fused_visc = !c->srcs.empty() && (c->srcs[0]->kind == MFT_SRC_UPWIND || c->srcs[0]->kind == MFT_SRC_RESIDUAL);
    const bool p2p_rv = c->p2p && c->nranks > 1 && fused_visc && c->srcs[0]->kind == MFT_SRC_RESIDUAL;
    if (p2p_rv) {
        // the norms only need OWNED points, so their two flag round-trips are interleaved with the halo put / wait:
        // sums fly while the halo block is being written, candidates fly while we wait for the neighbours' halo
        CHECK(p2p_norms_part(c, 0));
        CHECK(p2p_put<4>(c, c->u.p));
        CHECK(p2p_norms_part(c, 1));
        CHECK(p2p_wait(c, 0));
        CHECK(p2p_norms_part(c, 2));
    } else {
        NvtxRange r("update halos");
        CHECK(halo_exchange<4 /*V set below*/>(c, c->u.p));
    }
    size_t first = 0;
    if (fused_visc) {
        Source *s = c->srcs[0];
        const int visc = s->kind == MFT_SRC_UPWIND ? VISC_UPWIND : VISC_RESIDUAL;
        NvtxRange r(visc == VISC_RESIDUAL ? "calc fluxes + calc SourceResidualViscosityTominec (fused)"
                                          : "calc fluxes + calc SourceUpwindViscosityTominec (fused)");
        if (visc == VISC_RESIDUAL && !p2p_rv) CHECK(c->nranks > 1 ? launch_norms_multi(c) : launch_norms(c));
        CHECK(launch_pass_a(c, true, visc, s, false));  // flux divergence + D u + eps + g in one sweep
        CHECK(halo_exchange<8>(c, c->g.p));
        CHECK(launch_pass_b(c));
        first = 1;
    } else {
        NvtxRange r("calc fluxes");
        CHECK(launch_pass_a(c, true, VISC_NONE, nullptr, false));
    }
    {
        NvtxRange r("source terms");
        for (size_t i = first; i < c->srcs.size(); ++i) CHECK(apply_source_dev(c, c->srcs[i]));
    }
    NvtxRange r("boundary flux");
    CHECK(launch_boundary(c, true));  // pass 2
    return MFT_OK;
}

extern "C" int mft_rhs(mft_ctx *c, double t, double *const *u_soa, double *const *du_soa, int mem)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (c->V == 1 && c->nranks > 1) return fail(MFT_ENOTSUP, "multi-rank advection is not implemented");
    if (mem == MFT_MEM_HOST) {
        CHECK(upload_soa(c, u_soa, c->u.p));
        c->have_fsal = false;
    } else if (mem != MFT_MEM_DEVICE) {
        return fail(MFT_EINVAL, "mft_rhs: mem must be MFT_MEM_HOST or MFT_MEM_DEVICE");
    }
    CHECK(rhs_device(c, t));
    if (mem == MFT_MEM_HOST) {
        // rhs! changes u only at boundary points (strong BCs) and in the halo tail: bring back just those rows
        const int64_t m = (int64_t)c->touched_caller.size();
        if (m > 0) {
            ScopedTimer tm(c, MFT_K_OTHER);
            if (c->V == 4)
                k_gather_rows<4><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const Vec<4> *>(c->u.p), c->touched_dev.p, c->touched_buf.p, m);
            else
                k_gather_rows<1><<<grid_for(m, 256), 256, 0, c->stream>>>(reinterpret_cast<const Vec<1> *>(c->u.p), c->touched_dev.p, c->touched_buf.p, m);
            c->launches++;
            LAUNCH_CHECK();
            CU(cudaMemcpyAsync(c->touched_host, c->touched_buf.p, sizeof(double) * m * c->V, cudaMemcpyDeviceToHost, c->stream));
        }
        CHECK(download_soa(c, c->du.p, du_soa));
        CU(cudaStreamSynchronize(c->stream));
        for (int v = 0; v < c->V; ++v) {
            double *dst = u_soa[v];
            const double *src = c->touched_host + (int64_t)v * m;
            for (int64_t i = 0; i < m; ++i) dst[c->touched_caller[i]] = src[i];
        }
        return MFT_OK;
    }
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

extern "C" int mft_calc_fluxes(mft_ctx *c, double *const *u_soa, double *const *du_soa)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    CHECK(upload_soa(c, u_soa, c->u.p));
    CHECK(upload_soa(c, du_soa, c->du.p));
    c->have_fsal = false;
    CHECK(launch_pass_a(c, true, VISC_NONE, nullptr, true));
    CHECK(download_soa(c, c->du.p, du_soa));
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

extern "C" int mft_apply_source(mft_ctx *c, int index, double t, double *const *u_soa, double *const *du_soa)
{
    (void)t;
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (index < 0 || index >= (int)c->srcs.size()) return fail(MFT_EINVAL, "mft_apply_source: index %d out of range", index);
    CHECK(upload_soa(c, u_soa, c->u.p));
    CHECK(upload_soa(c, du_soa, c->du.p));
    c->have_fsal = false;
    CHECK(apply_source_dev(c, c->srcs[index]));
    CHECK(download_soa(c, c->du.p, du_soa));
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

extern "C" int mft_boundary_pass(mft_ctx *c, double t, double *const *u_soa, double *const *du_soa)
{
    (void)t;
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    CHECK(upload_soa(c, u_soa, c->u.p));
    CHECK(upload_soa(c, du_soa, c->du.p));
    c->have_fsal = false;
    CHECK(launch_boundary(c, true));
    CHECK(download_soa(c, c->u.p, u_soa));
    CU(cudaStreamSynchronize(c->stream));
    CHECK(download_soa(c, c->du.p, du_soa));
    CU(cudaStreamSynchronize(c->stream));
    return MFT_OK;
}

#include "mft_time_loop.inl"

extern "C" int mft_synchronize(mft_ctx *c)
{
    NEED_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    if (c->p2p) {
        P2PLocal h;
        CU(cudaMemcpy(&h, c->p2p_local.p, sizeof h, cudaMemcpyDeviceToHost));
        if (h.error) return fail(MFT_ENCCL, "peer-memory exchange timed out waiting for a peer flag (a rank fell out of step)");
    }
    return check_norm_misses(c);
}

extern "C" int mft_timer_start(mft_ctx *c)
{
    NEED_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventRecord(c->ev_t0, c->stream));
    return MFT_OK;
}

extern "C" int mft_timer_stop(mft_ctx *c, double *elapsed_ms)
{
    NEED_CTX(c);
    CU(cudaEventRecord(c->ev_t1, c->stream));
    CU(cudaEventSynchronize(c->ev_t1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    if (elapsed_ms) *elapsed_ms = (double)ms;
    return MFT_OK;
}

extern "C" int64_t mft_launch_count(mft_ctx *c) { return c ? c->launches : 0; }

extern "C" int mft_set_kernel_timing(mft_ctx *c, int enable)
{
    NEED_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    for (auto ev : c->kt.ev) cudaEventDestroy(ev);
    c->kt.ev.clear();
    c->kt.cls.clear();
    c->timing = enable != 0;
    return MFT_OK;
}

extern "C" int mft_kernel_time_ms(mft_ctx *c, int which, double *ms, int64_t *launches)
{
    NEED_CTX(c);
    CU(cudaStreamSynchronize(c->stream));
    if (which < 0) return mft_set_kernel_timing(c, c->timing);
    double tot = 0.0;
    int64_t cnt = 0;
    for (size_t i = 0; i < c->kt.cls.size(); ++i) {
        if (c->kt.cls[i] != which) continue;
        float e = 0.f;
        CU(cudaEventElapsedTime(&e, c->kt.ev[2 * i], c->kt.ev[2 * i + 1]));
        tot += e;
        cnt++;
    }
    if (ms) *ms = tot;
    if (launches) *launches = cnt;
    return MFT_OK;
}

extern "C" int mft_get_field(mft_ctx *c, int field, double *out)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (!out) return fail(MFT_EINVAL, "mft_get_field: out is NULL");
    const int64_t n = c->n_tot;
    CU(cudaStreamSynchronize(c->stream));
    auto scalar_ptr = [&](const double *bp) -> int {
        if (!bp) return fail(MFT_EINVAL, "mft_get_field: field not available (enable MFT_OPT_DIAGNOSTICS before the first compute call)");
        k_unpack_scalar<<<grid_for(n, 256), 256, 0, c->stream>>>(bp, c->d_perm.p, c->stage_soa.p, n);
        c->launches++;
        LAUNCH_CHECK();
        CU(cudaMemcpyAsync(out, c->stage_soa.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return MFT_OK;
    };
    auto scalar = [&](DevBuf<double> &b) -> int { return scalar_ptr(b.p); };
    auto vecfield = [&](DevBuf<double> &b) -> int {
        if (!b.p) return fail(MFT_EINVAL, "mft_get_field: field not available");
        std::vector<double *> ptrs(c->V);
        for (int v = 0; v < c->V; ++v) ptrs[v] = out + (int64_t)v * n;
        CHECK(download_soa(c, b.p, ptrs.data()));
        CU(cudaStreamSynchronize(c->stream));
        return MFT_OK;
    };
    switch (field) {
    case MFT_FIELD_EPS: return scalar(c->eps);
    case MFT_FIELD_EPS_UW: return scalar(c->eps_uw);
    case MFT_FIELD_EPS_RV: return scalar(c->eps_rv);
    case MFT_FIELD_EPS_C: return scalar(c->eps_c);
    case MFT_FIELD_RESIDUAL: return vecfield(c->residual);
    case MFT_FIELD_APPROX_DU: return vecfield(c->approx_du);
    case MFT_FIELD_SIGMA:
        for (auto *s : c->srcs)
            if (s->kind == MFT_SRC_IGR) return scalar_ptr(s->igr_args.x);
        return fail(MFT_EINVAL, "mft_get_field: no IGR source");
    case MFT_FIELD_IGR_STATUS: {
        for (auto *s : c->srcs)
            if (s->kind == MFT_SRC_IGR) {
                mft_igr::IgrScalars S;
                CU(cudaMemcpy(&S, s->igr_scalars.p, sizeof S, cudaMemcpyDeviceToHost));
                out[0] = (double)S.iter;
                out[1] = S.res;
                out[2] = S.res0;
                return MFT_OK;
            }
        return fail(MFT_EINVAL, "mft_get_field: no IGR source");
    }
    case MFT_FIELD_NORM_MISSES: {
        unsigned long long m = 0;
        CU(cudaMemcpy(&m, c->norm_miss.p, sizeof m, cudaMemcpyDeviceToHost));
        out[0] = (double)m;
        return MFT_OK;
    }
    case MFT_FIELD_NORMS:
        CU(cudaMemcpy(out, c->stats.p + 2 * c->V, sizeof(double) * c->V, cudaMemcpyDeviceToHost));
        return MFT_OK;
    default: return fail(MFT_EINVAL, "mft_get_field: unknown field %d", field);
    }
}

// failure detection: number of non-finite entries in the owned rows of the resident state
extern "C" int mft_count_nonfinite(mft_ctx *c, int64_t *count_out)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (!count_out) return fail(MFT_EINVAL, "mft_count_nonfinite: count_out is NULL");
    DevBuf<unsigned long long> cnt;
    CHECK(cnt.alloc(1));
    CU(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), c->stream));
    k_count_nonfinite<<<c->red_blocks, 256, 0, c->stream>>>(c->u.p, c->V, c->n_local, cnt.p);
    c->launches++;
    const cudaError_t le = cudaGetLastError();
    unsigned long long h = 0;
    cudaError_t ce = le == cudaSuccess ? cudaMemcpyAsync(&h, cnt.p, sizeof h, cudaMemcpyDeviceToHost, c->stream) : le;
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    cnt.release();
    if (ce != cudaSuccess) return fail(MFT_ECUDA, "mft_count_nonfinite: %s", cudaGetErrorString(ce));
    *count_out = (int64_t)h;
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// pinned host memory helpers
// ------------------------------------------------------------------------------------------------------
extern "C" int mft_host_alloc(void **out, int64_t bytes)
{
    if (!out || bytes <= 0) return fail(MFT_EINVAL, "mft_host_alloc: bad arguments");
    CU(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
    return MFT_OK;
}
extern "C" int mft_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return MFT_OK;
}
extern "C" int mft_host_register(void *p, int64_t bytes)
{
    if (!p || bytes <= 0) return fail(MFT_EINVAL, "mft_host_register: bad arguments");
    CU(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
    return MFT_OK;
}
extern "C" int mft_host_unregister(void *p)
{
    if (p) CU(cudaHostUnregister(p));
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// Hilbert space-filling-curve ordering (setup helper, host only)
// ------------------------------------------------------------------------------------------------------
static inline uint64_t hilbert_d(uint32_t x, uint32_t y, int bits)
{
    uint64_t d = 0;
    for (uint32_t s = 1u << (bits - 1); s > 0; s >>= 1) {
        const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
        d += (uint64_t)s * s * ((3u * rx) ^ ry);
        if (ry == 0) {
            if (rx == 1) {
                x = (1u << bits) - 1u - x;
                y = (1u << bits) - 1u - y;
            }
            const uint32_t tmp = x;
            x = y;
            y = tmp;
        }
    }
    return d;
}

#include "mft_setup_cuda.inl"

extern "C" int mft_sfc_order(int64_t n, const double *x, const double *y, int64_t *perm1_out)
{
    if (n <= 0 || !x || !y || !perm1_out) return fail(MFT_EINVAL, "mft_sfc_order: bad arguments");
    double xmin = x[0], xmax = x[0], ymin = y[0], ymax = y[0];
    for (int64_t i = 1; i < n; ++i) {
        xmin = std::min(xmin, x[i]);
        xmax = std::max(xmax, x[i]);
        ymin = std::min(ymin, y[i]);
        ymax = std::max(ymax, y[i]);
    }
    const int bits = 20;
    const double ext = std::max(std::max(xmax - xmin, ymax - ymin), 1e-300);
    const double scale = ((double)((1u << bits) - 1u)) / ext;
    std::vector<std::pair<uint64_t, int64_t>> keyed((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t qx = (uint32_t)((x[i] - xmin) * scale), qy = (uint32_t)((y[i] - ymin) * scale);
        keyed[i] = {hilbert_d(qx, qy, bits), i};
    }
    std::sort(keyed.begin(), keyed.end());
    for (int64_t d = 0; d < n; ++d) perm1_out[d] = keyed[d].second + 1;
    return MFT_OK;
}

#include "mft_multi_gpu.inl"
