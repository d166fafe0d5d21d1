// psb_comm.cu -- multi-GPU plumbing owned by the library: NCCL (loaded with dlopen, no link or
// header dependency) for the one collective of the path, the gather of the per-variant result
// table, plus the small broadcast / all-reduce / barrier a launcher needs around it.
//
// Reference being replaced: multiprocessing.Pool(options.cpu) and the ordered pool.starmap of
// pyseer/__main__.py:517-519, :541-546, :777-780 -- variants are independent given the once-per-run
// state, so ranks take contiguous variant ranges and the only exchange is on the way out.
//
// A psb_comm holds the LOCAL members of a communicator: one context when every GPU has its own
// process (psb_comm_init_rank; torchrun-style launches), all of them when one process drives
// several GPUs (psb_comm_init_all; the CLI's --gpus N).  Every operation loops over the local
// members inside one NCCL group.
//
// Gather pipeline per member (psb_comm_gather_begin):
//   compute stream --ev_ready--> comm stream: pack the table (device-to-device, one header + the
//   columns at fixed offsets for rows_max rows) --ev_packed--> compute stream (the next run may
//   overwrite the table) ; comm stream: ncclSend to the root, the root's ncclRecv from every rank.
// The NCCL transfer of step i therefore overlaps the kernels of step i + 1; psb_comm_gather_wait
// joins the comm stream back into the compute stream (device side) and the host.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "psb_internal.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return PSB_OK;
    const char *env = getenv("PSB_NCCL_LIB");
    void *h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    PSB_REQUIRE(h, PSB_ERR_UNSUPPORTED, "NCCL not found (dlopen libnccl.so.2: %s); set PSB_NCCL_LIB", dlerror());
#define SYM(field, name)                                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                            \
    PSB_REQUIRE(g_nccl.field, PSB_ERR_UNSUPPORTED, "NCCL symbol %s missing", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
    g_nccl.handle = h;
    return PSB_OK;
}

#define PSB_NCCL(call)                                                                     \
    do {                                                                                   \
        int r_ = (call);                                                                   \
        if (r_ != 0) {                                                                     \
            psb_set_error("%s:%d NCCL error %d: %s", __FILE__, __LINE__, r_,               \
                          g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");        \
            return PSB_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

// packed table of one rank: header, then the columns of psb_results for rows_max rows
struct CommHeader {
    int64_t n_rows;
    int64_t counts[4];      // loaded, pre-filtered, tested, passed (psb_counts)
    int64_t stats[3];       // psb_last_stats
};
static_assert(sizeof(CommHeader) == 64, "header is 64 bytes");

struct CommLayout {
    int64_t rows_max = 0;
    int nb = 0;             // slope columns (fixed effects)
    size_t off_carriers, off_missing, off_af, off_prep, off_pvalue, off_beta, off_bse, off_extra,
        off_flags, off_betas, bytes;
    void set(int64_t rows, int nbetas) {
        rows_max = rows;
        nb = nbetas;
        size_t o = sizeof(CommHeader);
        const size_t r = (size_t)((rows + 1) / 2 * 2);      // keep the fp64 columns 8-byte aligned
        off_carriers = o; o += r * 4;
        off_missing = o; o += r * 4;
        off_af = o; o += r * 8;
        off_prep = o; o += r * 8;
        off_pvalue = o; o += r * 8;
        off_beta = o; o += r * 8;
        off_bse = o; o += r * 8;
        off_extra = o; o += r * 8;
        off_flags = o; o += r * 4;
        off_betas = o; o += r * 8 * (size_t)nb;
        bytes = (o + 15) / 16 * 16;
    }
};

struct CommMember {
    psb_ctx *ctx = nullptr;
    int rank = 0;
    ncclComm_t nccl = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_packed = nullptr, ev_done = nullptr;
    uint8_t *d_pack = nullptr;      // this rank's packed table
    size_t pack_cap = 0;
    uint8_t *d_recv = nullptr;      // root: world x layout.bytes
    size_t recv_cap = 0;
    uint8_t *d_small = nullptr;     // scratch for bcast / allreduce
    size_t small_cap = 0;
};

struct psb_comm {
    int world = 1;
    std::vector<CommMember> m;
    CommLayout lay;
    int root = 0;
    bool gathered = false;
};

__global__ void k_comm_header(CommHeader *h, const int *__restrict__ counters, int64_t S) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        h->n_rows = S;
        h->counts[0] = S;
        h->counts[1] = counters[1] + counters[5];
        h->counts[2] = counters[0] - counters[5];
        h->counts[3] = counters[0] - counters[5] - counters[2];
        h->stats[0] = counters[4];
        h->stats[1] = counters[3];
        h->stats[2] = counters[2];
    }
}

static int reserve(uint8_t **p, size_t *cap, size_t bytes, cudaStream_t st) {
    if (bytes <= *cap) return PSB_OK;
    PSB_CUDA(cudaStreamSynchronize(st));
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    PSB_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    return PSB_OK;
}

static int member_init(CommMember &mb) {
    PSB_CUDA(cudaSetDevice(mb.ctx->device));
    PSB_CUDA(cudaStreamCreateWithFlags(&mb.stream, cudaStreamNonBlocking));
    PSB_CUDA(cudaEventCreateWithFlags(&mb.ev_ready, cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&mb.ev_packed, cudaEventDisableTiming));
    PSB_CUDA(cudaEventCreateWithFlags(&mb.ev_done, cudaEventDisableTiming));
    return PSB_OK;
}

extern "C" {

int psb_comm_unique_id(uint8_t id[PSB_COMM_ID_BYTES]) {
    PSB_REQUIRE(id, PSB_ERR_ARG, "id is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    PSB_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, PSB_COMM_ID_BYTES);
    return PSB_OK;
}

int psb_comm_init_rank(psb_ctx *ctx, int32_t world, int32_t rank, const uint8_t id[PSB_COMM_ID_BYTES],
                       psb_comm **out) {
    PSB_REQUIRE(ctx && id && out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(world >= 1 && rank >= 0 && rank < world, PSB_ERR_ARG, "bad rank %d of %d", rank, world);
    *out = nullptr;
    int rc = nccl_load();
    if (rc) return rc;
    psb_comm *c = new psb_comm();
    c->world = world;
    c->m.resize(1);
    c->m[0].ctx = ctx;
    c->m[0].rank = rank;
    rc = member_init(c->m[0]);
    if (rc) { delete c; return rc; }
    ncclUniqueId u;
    memcpy(u.internal, id, PSB_COMM_ID_BYTES);
    PSB_NCCL(g_nccl.CommInitRank(&c->m[0].nccl, world, u, rank));
    *out = c;
    return PSB_OK;
}

int psb_comm_init_all(psb_ctx *const *ctxs, int32_t n, psb_comm **out) {
    PSB_REQUIRE(ctxs && out && n >= 1, PSB_ERR_ARG, "bad arguments");
    *out = nullptr;
    int rc = nccl_load();
    if (rc) return rc;
    psb_comm *c = new psb_comm();
    c->world = n;
    c->m.resize(n);
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        PSB_REQUIRE(ctxs[i], PSB_ERR_ARG, "ctx %d is NULL", i);
        for (int j = 0; j < i; ++j)
            PSB_REQUIRE(ctxs[j]->device != ctxs[i]->device, PSB_ERR_ARG,
                        "contexts %d and %d share device %d", j, i, ctxs[i]->device);
        c->m[i].ctx = ctxs[i];
        c->m[i].rank = i;
        devs[i] = ctxs[i]->device;
        rc = member_init(c->m[i]);
        if (rc) { delete c; return rc; }
    }
    std::vector<ncclComm_t> comms(n);
    PSB_NCCL(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) c->m[i].nccl = comms[i];
    *out = c;
    return PSB_OK;
}

int psb_comm_destroy(psb_comm *c) {
    if (!c) return PSB_OK;
    for (auto &mb : c->m) {
        cudaSetDevice(mb.ctx->device);
        if (mb.stream) cudaStreamSynchronize(mb.stream);
        if (mb.nccl) g_nccl.CommDestroy(mb.nccl);
        if (mb.d_pack) cudaFree(mb.d_pack);
        if (mb.d_recv) cudaFree(mb.d_recv);
        if (mb.d_small) cudaFree(mb.d_small);
        if (mb.ev_ready) cudaEventDestroy(mb.ev_ready);
        if (mb.ev_packed) cudaEventDestroy(mb.ev_packed);
        if (mb.ev_done) cudaEventDestroy(mb.ev_done);
        if (mb.stream) cudaStreamDestroy(mb.stream);
    }
    delete c;
    return PSB_OK;
}

int psb_comm_info(psb_comm *c, int32_t *world, int32_t *n_local, int32_t *nccl_version) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "comm is NULL");
    if (world) *world = c->world;
    if (n_local) *n_local = (int32_t)c->m.size();
    if (nccl_version) {
        int v = 0;
        g_nccl.GetVersion(&v);
        *nccl_version = v;
    }
    return PSB_OK;
}

// ---- small collectives on host buffers (one local member: the one-process-per-GPU launch) ----
static int small_reserve(CommMember &mb, size_t bytes) {
    return reserve(&mb.d_small, &mb.small_cap, bytes < 256 ? 256 : bytes, mb.stream);
}

int psb_comm_bcast(psb_comm *c, void *host_buf, size_t bytes, int32_t root) {
    PSB_REQUIRE(c && (host_buf || bytes == 0), PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(c->m.size() == 1, PSB_ERR_UNSUPPORTED, "psb_comm_bcast needs one context per process");
    PSB_REQUIRE(root >= 0 && root < c->world, PSB_ERR_ARG, "bad root");
    if (bytes == 0) return PSB_OK;
    CommMember &mb = c->m[0];
    PSB_CUDA(cudaSetDevice(mb.ctx->device));
    int rc = small_reserve(mb, bytes);
    if (rc) return rc;
    if (mb.rank == root)
        PSB_CUDA(cudaMemcpyAsync(mb.d_small, host_buf, bytes, cudaMemcpyHostToDevice, mb.stream));
    PSB_NCCL(g_nccl.Broadcast(mb.d_small, mb.d_small, bytes, NCCL_UINT8, root, mb.nccl, mb.stream));
    if (mb.rank != root)
        PSB_CUDA(cudaMemcpyAsync(host_buf, mb.d_small, bytes, cudaMemcpyDeviceToHost, mb.stream));
    PSB_CUDA(cudaStreamSynchronize(mb.stream));
    if (bytes > (64u << 20)) {          // do not keep a large staging buffer around
        cudaFree(mb.d_small);
        mb.d_small = nullptr;
        mb.small_cap = 0;
    }
    return PSB_OK;
}

int psb_comm_allreduce(psb_comm *c, double *vals, int32_t n, int32_t op) {
    PSB_REQUIRE(c && vals && n > 0, PSB_ERR_ARG, "bad arguments");
    PSB_REQUIRE(c->m.size() == 1, PSB_ERR_UNSUPPORTED, "psb_comm_allreduce needs one context per process");
    PSB_REQUIRE(op == 0 || op == 1, PSB_ERR_ARG, "op: 0 = sum, 1 = max");
    CommMember &mb = c->m[0];
    PSB_CUDA(cudaSetDevice(mb.ctx->device));
    int rc = small_reserve(mb, (size_t)n * sizeof(double));
    if (rc) return rc;
    PSB_CUDA(cudaMemcpyAsync(mb.d_small, vals, n * sizeof(double), cudaMemcpyHostToDevice, mb.stream));
    PSB_NCCL(g_nccl.AllReduce(mb.d_small, mb.d_small, n, NCCL_FLOAT64, op == 0 ? NCCL_SUM : NCCL_MAX, mb.nccl,
                              mb.stream));
    PSB_CUDA(cudaMemcpyAsync(vals, mb.d_small, n * sizeof(double), cudaMemcpyDeviceToHost, mb.stream));
    PSB_CUDA(cudaStreamSynchronize(mb.stream));
    return PSB_OK;
}

int psb_comm_barrier(psb_comm *c) {
    PSB_REQUIRE(c, PSB_ERR_ARG, "comm is NULL");
    if (c->m.size() != 1) {             // one process: nothing to meet but the streams
        for (auto &mb : c->m) {
            PSB_CUDA(cudaSetDevice(mb.ctx->device));
            PSB_CUDA(cudaStreamSynchronize(mb.ctx->stream));
            PSB_CUDA(cudaStreamSynchronize(mb.stream));
        }
        return PSB_OK;
    }
    PSB_CUDA(cudaSetDevice(c->m[0].ctx->device));
    PSB_CUDA(cudaStreamSynchronize(c->m[0].ctx->stream));
    double one = 1.0;
    return psb_comm_allreduce(c, &one, 1, 0);
}

// ---- the gather of the result table ----
int psb_comm_gather_begin(psb_comm *c, int32_t root, int64_t rows_max) {
    PSB_NVTX("psb_comm_gather_begin");
    PSB_REQUIRE(c, PSB_ERR_ARG, "comm is NULL");
    PSB_REQUIRE(root >= 0 && root < c->world && rows_max >= 0, PSB_ERR_ARG, "bad root / rows_max");
    int nb = 0;
    for (auto &mb : c->m) {
        psb_ctx *x = mb.ctx;
        PSB_REQUIRE(x->ran, PSB_ERR_STATE, "psb_comm_gather_begin before psb_run_*");
        PSB_REQUIRE(x->S <= rows_max, PSB_ERR_ARG, "rank %d holds %lld rows, rows_max is %lld", mb.rank,
                    (long long)x->S, (long long)rows_max);
        nb = (x->model == PSB_MODEL_FIXED && x->q > 1) ? x->q - 1 : 0;
    }
    if (c->lay.rows_max != rows_max || c->lay.nb != nb) c->lay.set(rows_max, nb);
    const CommLayout &L = c->lay;
    c->root = root;
    // 1. pack (device to device) behind the run, release the table to the next run
    for (auto &mb : c->m) {
        psb_ctx *x = mb.ctx;
        PSB_CUDA(cudaSetDevice(x->device));
        int rc = reserve(&mb.d_pack, &mb.pack_cap, L.bytes, mb.stream);
        if (rc) return rc;
        if (mb.rank == root) {
            rc = reserve(&mb.d_recv, &mb.recv_cap, L.bytes * (size_t)c->world, mb.stream);
            if (rc) return rc;
        }
        PSB_CUDA(cudaEventRecord(mb.ev_ready, x->stream));
        PSB_CUDA(cudaStreamWaitEvent(mb.stream, mb.ev_ready, 0));
        k_comm_header<<<1, 32, 0, mb.stream>>>((CommHeader *)mb.d_pack, x->d_counters, x->S);
        PSB_CUDA(cudaGetLastError());
        x->launches++;
        const int64_t S = x->S;
#define PK(off, src, type)                                                                            \
    if (S > 0)                                                                                        \
        PSB_CUDA(cudaMemcpyAsync(mb.d_pack + L.off, src, (size_t)S * sizeof(type), cudaMemcpyDeviceToDevice, \
                                 mb.stream));
        PK(off_carriers, x->d_carriers, int32_t)
        PK(off_missing, x->d_missing, int32_t)
        PK(off_af, x->d_af, double)
        PK(off_prep, x->d_prep, double)
        PK(off_pvalue, x->d_pvalue, double)
        PK(off_beta, x->d_beta, double)
        PK(off_bse, x->d_bse, double)
        PK(off_extra, x->d_extra, double)
        PK(off_flags, x->d_flags, uint32_t)
#undef PK
        if (S > 0 && nb > 0)
            PSB_CUDA(cudaMemcpyAsync(mb.d_pack + L.off_betas, x->d_betas, (size_t)S * nb * sizeof(double),
                                     cudaMemcpyDeviceToDevice, mb.stream));
        PSB_CUDA(cudaEventRecord(mb.ev_packed, mb.stream));
        PSB_CUDA(cudaStreamWaitEvent(x->stream, mb.ev_packed, 0));
    }
    // 2. one NCCL group: every rank sends its packed table, the root receives world of them
    PSB_NCCL(g_nccl.GroupStart());
    for (auto &mb : c->m) {
        PSB_NCCL(g_nccl.Send(mb.d_pack, L.bytes, NCCL_UINT8, root, mb.nccl, mb.stream));
        if (mb.rank == root)
            for (int r = 0; r < c->world; ++r)
                PSB_NCCL(g_nccl.Recv(mb.d_recv + (size_t)r * L.bytes, L.bytes, NCCL_UINT8, r, mb.nccl, mb.stream));
    }
    PSB_NCCL(g_nccl.GroupEnd());
    c->gathered = true;
    return PSB_OK;
}

int psb_comm_gather_wait(psb_comm *c) {
    PSB_NVTX("psb_comm_gather_wait");
    PSB_REQUIRE(c, PSB_ERR_ARG, "comm is NULL");
    for (auto &mb : c->m) {
        PSB_CUDA(cudaSetDevice(mb.ctx->device));
        PSB_CUDA(cudaEventRecord(mb.ev_done, mb.stream));
        PSB_CUDA(cudaStreamWaitEvent(mb.ctx->stream, mb.ev_done, 0));    // events recorded next on the
    }                                                                     // compute stream include the gather
    for (auto &mb : c->m) {
        PSB_CUDA(cudaSetDevice(mb.ctx->device));
        PSB_CUDA(cudaStreamSynchronize(mb.stream));
    }
    return PSB_OK;
}

int psb_comm_gather_fetch(psb_comm *c, int32_t src_rank, const psb_results *out, int64_t *n_rows,
                          int64_t counts[4]) {
    PSB_REQUIRE(c && src_rank >= 0 && src_rank < c->world, PSB_ERR_ARG, "bad arguments");
    PSB_REQUIRE(c->gathered, PSB_ERR_STATE, "psb_comm_gather_fetch before psb_comm_gather_begin");
    CommMember *rootm = nullptr;
    for (auto &mb : c->m)
        if (mb.rank == c->root) rootm = &mb;
    PSB_REQUIRE(rootm, PSB_ERR_STATE, "the root rank %d is not in this process", c->root);
    const CommLayout &L = c->lay;
    PSB_CUDA(cudaSetDevice(rootm->ctx->device));
    PSB_CUDA(cudaStreamSynchronize(rootm->stream));
    const uint8_t *base = rootm->d_recv + (size_t)src_rank * L.bytes;
    CommHeader h;
    PSB_CUDA(cudaMemcpy(&h, base, sizeof(h), cudaMemcpyDeviceToHost));
    PSB_REQUIRE(h.n_rows >= 0 && h.n_rows <= L.rows_max, PSB_ERR_STATE, "corrupt table header from rank %d", src_rank);
    if (n_rows) *n_rows = h.n_rows;
    if (counts) memcpy(counts, h.counts, sizeof(h.counts));
    const int64_t S = h.n_rows;
    if (out && S > 0) {
        cudaStream_t st = rootm->stream;
#define CP(field, off, type)                                                                          \
    if (out->field) PSB_CUDA(cudaMemcpyAsync(out->field, base + L.off, (size_t)S * sizeof(type), cudaMemcpyDefault, st));
        CP(carriers, off_carriers, int32_t)
        CP(missing, off_missing, int32_t)
        CP(af, off_af, double)
        CP(prep, off_prep, double)
        CP(pvalue, off_pvalue, double)
        CP(beta, off_beta, double)
        CP(bse, off_bse, double)
        CP(extra, off_extra, double)
        CP(flags, off_flags, uint32_t)
#undef CP
        if (out->betas && L.nb > 0)
            PSB_CUDA(cudaMemcpyAsync(out->betas, base + L.off_betas, (size_t)S * L.nb * sizeof(double),
                                     cudaMemcpyDefault, st));
        PSB_CUDA(cudaStreamSynchronize(st));
    }
    return PSB_OK;
}

int psb_comm_gather_bytes(psb_comm *c, int64_t *bytes_per_rank) {
    PSB_REQUIRE(c && bytes_per_rank, PSB_ERR_ARG, "NULL argument");
    *bytes_per_rank = (int64_t)c->lay.bytes;
    return PSB_OK;
}

}  // extern "C"
