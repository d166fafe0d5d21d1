// psb_eigh.cu -- once-per-run symmetric eigendecomposition on the device.
//
// Replaces the host `eigh` of LMM.setSU_fromK (fastlmm/lmm_cov.py:88-103: K_ = P (K + I) P,
// S, U = eigh(K_)), the O(N^3) part of lmm.initialise_lmm (lmm.py:26-122): 4 s at N = 5000 and
// 23 s at N = 10 000 in NumPy on the GPU box's host, once per run.  This is plain library work
// (cuSOLVER's divide-and-conquer syevd in fp64), not a hand-written kernel: the library is
// loaded lazily with dlopen so that libpyseer_b200.so itself carries no link dependency on it,
// and PSB_ERR_UNSUPPORTED tells the caller (pyseer_b200/lmm.py) to keep its NumPy eigh.
#include <dlfcn.h>

#include "psb_internal.cuh"

namespace {

typedef void *solver_handle;
typedef int (*fn_create)(solver_handle *);
typedef int (*fn_destroy)(solver_handle);
typedef int (*fn_set_stream)(solver_handle, cudaStream_t);
typedef int (*fn_syevd_bufsize)(solver_handle, int jobz, int uplo, int n, const double *A, int lda,
                                const double *W, int *lwork);
typedef int (*fn_syevd)(solver_handle, int jobz, int uplo, int n, double *A, int lda, double *W,
                        double *work, int lwork, int *info);

struct Solver {
    void *lib = nullptr;
    fn_create create = nullptr;
    fn_destroy destroy = nullptr;
    fn_set_stream set_stream = nullptr;
    fn_syevd_bufsize bufsize = nullptr;
    fn_syevd syevd = nullptr;
    bool tried = false;
};
Solver g_solver;

bool load_solver() {
    if (g_solver.tried) return g_solver.syevd != nullptr;
    g_solver.tried = true;
    const char *names[] = {"libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"};
    for (const char *nm : names) {
        g_solver.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (g_solver.lib) break;
    }
    if (!g_solver.lib) return false;
    g_solver.create = (fn_create)dlsym(g_solver.lib, "cusolverDnCreate");
    g_solver.destroy = (fn_destroy)dlsym(g_solver.lib, "cusolverDnDestroy");
    g_solver.set_stream = (fn_set_stream)dlsym(g_solver.lib, "cusolverDnSetStream");
    g_solver.bufsize = (fn_syevd_bufsize)dlsym(g_solver.lib, "cusolverDnDsyevd_bufferSize");
    g_solver.syevd = (fn_syevd)dlsym(g_solver.lib, "cusolverDnDsyevd");
    if (!(g_solver.create && g_solver.destroy && g_solver.set_stream && g_solver.bufsize && g_solver.syevd)) {
        g_solver.syevd = nullptr;
        return false;
    }
    return true;
}

// out[i][j] = in[j][i], n x n, 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
k_transpose(const double *__restrict__ in, double *__restrict__ out, int n) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i < n && j < n) t[r][threadIdx.x] = in[(size_t)i * n + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = bx + r, j = by + threadIdx.x;
        if (i < n && j < n) out[(size_t)i * n + j] = t[threadIdx.x][r];
    }
}

// T[k][j] (+)= sum_i Xp[k][i] A[i][j] over this block's slice of i   (T zeroed by the caller)
__global__ void __launch_bounds__(256)
k_proj_T(const double *__restrict__ A, const double *__restrict__ Xp, double *__restrict__ T, int n, int d,
         int rows_per_block) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * rows_per_block, i1 = min(n, i0 + rows_per_block);
    if (j >= n) return;
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int i = i0; i < i1; ++i) {
        const double a = A[(size_t)i * n + j];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < d) acc[k] = fma(__ldg(Xp + (size_t)k * n + i), a, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < d) atomicAdd(T + (size_t)k * n + j, acc[k]);
}

// T[k][j] = sum_i Xp[k][i] A[j][i]   (the same for the transposed matrix): one warp per row j
__global__ void __launch_bounds__(256)
k_proj_Tt(const double *__restrict__ A, const double *__restrict__ Xp, double *__restrict__ T, int n, int d) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double a = A[(size_t)j * n + i];
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < d) acc[k] = fma(__ldg(Xp + (size_t)k * n + i), a, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k < d) {
            double v = acc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) T[(size_t)k * n + j] = v;
        }
}

// out[i][j] = (TR ? A[j][i] : A[i][j]) - sum_k X[i][k] T[k][j]      (Linreg.regress, lmm_cov.py:874-880)
template <bool TR>
__global__ void __launch_bounds__(256)
k_proj_apply(const double *__restrict__ A, const double *__restrict__ X, const double *__restrict__ T,
             double *__restrict__ out, int n, int d, double diag_add) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    if (TR) {
        for (int r = threadIdx.y; r < 32; r += 8) {
            const int i = bx + r, j = by + threadIdx.x;          // read A[j'][i'] tile transposed
            if (i < n && j < n) t[r][threadIdx.x] = A[(size_t)i * n + j];
        }
        __syncthreads();
    }
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = by + r, j = bx + threadIdx.x;
        if (i < n && j < n) {
            double v = TR ? t[threadIdx.x][r] : A[(size_t)i * n + j];
            if (!TR && i == j) v += diag_add;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(__ldg(X + (size_t)i * d + k), __ldg(T + (size_t)k * n + j), s);
            out[(size_t)i * n + j] = v - s;
        }
    }
}

__global__ void k_add_diag(double *A, int n, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(size_t)i * n + i] += v;
}

}  // namespace

static int eigh_core(psb_ctx *c, int32_t n, const double *A, const double *X, const double *Xp, int32_t d,
                     double *w_out, double *V_out);

// A: n x n row-major, host; only its lower triangle is read (as numpy.linalg.eigh does).  w_out: n eigenvalues,
// ascending.  V_out: n x n row-major, column j = eigenvector of w_out[j] (numpy.linalg.eigh layout).
extern "C" int psb_eigh(psb_ctx *c, int32_t n, const double *A, double *w_out, double *V_out) {
    return eigh_core(c, n, A, nullptr, nullptr, 0, w_out, V_out);
}

// LMM.setSU_fromK (fastlmm/lmm_cov.py:88-103) in one call: K (n x n row-major, host) and the covariate
// design X (n x d) with its pseudo-inverse Xp (d x n, Linreg's Xdagger) -> eigendecomposition of
// K_ = P (K + I) P formed on the device exactly as the reference forms it (regress the matrix, then
// regress its transpose: two rank-d updates), eigenvalues ascending in w_out, eigenvectors in the
// columns of V_out (numpy.linalg.eigh layout).  The caller drops the first d pairs and subtracts 1.
extern "C" int psb_spectral(psb_ctx *c, int32_t n, int32_t d, const double *K, const double *X,
                            const double *Xp, double *w_out, double *V_out) {
    PSB_REQUIRE(X && Xp, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(d >= 1 && d <= 16, PSB_ERR_UNSUPPORTED, "psb_spectral supports 1..16 covariate columns, got %d", d);
    return eigh_core(c, n, K, X, Xp, d, w_out, V_out);
}

static int eigh_core(psb_ctx *c, int32_t n, const double *A, const double *X, const double *Xp, int32_t d,
                     double *w_out, double *V_out) {
    PSB_REQUIRE(c && A && w_out && V_out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(n >= 1 && n <= 46000, PSB_ERR_ARG, "n = %d out of range", n);
    PSB_REQUIRE(load_solver(), PSB_ERR_UNSUPPORTED,
                "cuSOLVER (libcusolver.so.11) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
    PSB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n * n * sizeof(double);
    double *d_A = nullptr, *d_V = nullptr, *d_w = nullptr, *d_work = nullptr;
    double *d_X = nullptr, *d_Xp = nullptr, *d_T = nullptr;
    int *d_info = nullptr;
    solver_handle h = nullptr;
    int rc = PSB_OK, info = 0, lwork = 0, st = 0;
#define EIGH_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            psb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e_), \
                          cudaGetErrorString(e_));                                        \
            rc = PSB_ERR_CUDA;                                                            \
            goto done;                                                                    \
        }                                                                                 \
    } while (0)
    EIGH_CUDA(cudaMalloc(&d_A, bytes));
    EIGH_CUDA(cudaMalloc(&d_V, bytes));
    EIGH_CUDA(cudaMalloc(&d_w, (size_t)n * sizeof(double)));
    EIGH_CUDA(cudaMalloc(&d_info, sizeof(int)));
    EIGH_CUDA(cudaMemcpyAsync(d_A, A, bytes, cudaMemcpyHostToDevice, c->stream));
    if (X) {
        // K_ = regress(regress(K + I)')  -- K_ = self.regress(self.K); K_ = self.regress(K_.T)
        const size_t xb = (size_t)n * d * sizeof(double);
        EIGH_CUDA(cudaMalloc(&d_X, xb));
        EIGH_CUDA(cudaMalloc(&d_Xp, xb));
        EIGH_CUDA(cudaMalloc(&d_T, xb));
        EIGH_CUDA(cudaMemcpyAsync(d_X, X, xb, cudaMemcpyHostToDevice, c->stream));
        EIGH_CUDA(cudaMemcpyAsync(d_Xp, Xp, xb, cudaMemcpyHostToDevice, c->stream));
        k_add_diag<<<(n + 255) / 256, 256, 0, c->stream>>>(d_A, n, 1.0);
        EIGH_CUDA(cudaMemsetAsync(d_T, 0, xb, c->stream));
        const int rpb = 256;
        k_proj_T<<<dim3((n + 255) / 256, (n + rpb - 1) / rpb), 256, 0, c->stream>>>(d_A, d_Xp, d_T, n, d, rpb);
        dim3 g((n + 31) / 32, (n + 31) / 32), b(32, 8);
        k_proj_apply<false><<<g, b, 0, c->stream>>>(d_A, d_X, d_T, d_V, n, d, 0.0);          // d_V = regress(K + I)
        k_proj_Tt<<<(n + 7) / 8, 256, 0, c->stream>>>(d_V, d_Xp, d_T, n, d);
        k_proj_apply<true><<<g, b, 0, c->stream>>>(d_V, d_X, d_T, d_A, n, d, 0.0);           // d_A = regress(d_V')
        c->launches += 5;
        EIGH_CUDA(cudaGetLastError());
    }
    st = g_solver.create(&h);
    if (st != 0) { psb_set_error("cusolverDnCreate failed (%d)", st); rc = PSB_ERR_CUDA; goto done; }
    g_solver.set_stream(h, c->stream);
    // jobz = CUSOLVER_EIG_MODE_VECTOR (1); uplo = CUBLAS_FILL_MODE_UPPER (1): the column-major
    // upper triangle is the row-major LOWER triangle, the half numpy.linalg.eigh reads
    st = g_solver.bufsize(h, 1, 1, n, d_A, n, d_w, &lwork);
    if (st != 0) { psb_set_error("cusolverDnDsyevd_bufferSize failed (%d)", st); rc = PSB_ERR_CUDA; goto done; }
    EIGH_CUDA(cudaMalloc(&d_work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double)));
    st = g_solver.syevd(h, 1, 1, n, d_A, n, d_w, d_work, lwork, d_info);
    if (st != 0) { psb_set_error("cusolverDnDsyevd failed (%d)", st); rc = PSB_ERR_CUDA; goto done; }
    EIGH_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    // cuSOLVER is column-major: eigenvector j sits in column j = d_A[j * n + i]; the caller wants
    // the row-major matrix with eigenvectors in columns
    {
        dim3 g((n + 31) / 32, (n + 31) / 32), b(32, 8);
        k_transpose<<<g, b, 0, c->stream>>>(d_A, d_V, n);
        c->launches++;
    }
    EIGH_CUDA(cudaGetLastError());
    EIGH_CUDA(cudaMemcpyAsync(V_out, d_V, bytes, cudaMemcpyDeviceToHost, c->stream));
    EIGH_CUDA(cudaMemcpyAsync(w_out, d_w, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    EIGH_CUDA(cudaStreamSynchronize(c->stream));
    if (info != 0) {
        psb_set_error("syevd did not converge (info = %d)", info);
        rc = PSB_ERR_NUMERIC;
    }
done:
#undef EIGH_CUDA
    if (h) g_solver.destroy(h);
    if (d_work) cudaFree(d_work);
    if (d_info) cudaFree(d_info);
    if (d_X) cudaFree(d_X);
    if (d_Xp) cudaFree(d_Xp);
    if (d_T) cudaFree(d_T);
    if (d_w) cudaFree(d_w);
    if (d_V) cudaFree(d_V);
    if (d_A) cudaFree(d_A);
    return rc;
}

// ---------------------------------------------------------------------------------------
// The h2 search of LMM.findH2 (fastlmm/lmm_cov.py:427-478 through mingrid.minimize1D,
// mingrid.py:13-73) evaluates the null-model likelihood nLLeval(h2) (lmm_cov.py:597-684) ~30 times;
// with the rotated phenotype U'Py fixed, each evaluation is two sums over the J = N - D spectrum:
//     yKy(h2) = sum_j uy_j^2 / (h2 S_j + 1 - h2),    logdetK(h2) = sum_j log(h2 S_j + 1 - h2).
// One block per requested h2 (the grid of the search goes in one launch, Brent's points one by
// one); fixed reduction tree, so the value of a given h2 is reproducible.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_nll_terms(const double *__restrict__ S, const double *__restrict__ uy2, int J,
            const double *__restrict__ h2, double *__restrict__ out) {
    __shared__ double red[2][8];
    const double h = h2[blockIdx.x];
    double a = 0.0, b = 0.0;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
        const double sd = h * S[j] + (1.0 - h);
        a += uy2[j] / sd;
        b += log(sd);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a;
        red[1][threadIdx.x >> 5] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int w = 0; w < 8; ++w) {
            ta += red[0][w];
            tb += red[1][w];
        }
        out[2 * blockIdx.x] = ta;
        out[2 * blockIdx.x + 1] = tb;
    }
}

extern "C" int psb_lmm_nll_terms(psb_ctx *c, int32_t J, const double *S, const double *uy2, int32_t n_h2,
                                 const double *h2, double *yky_out, double *logdet_out) {
    PSB_REQUIRE(c && S && uy2 && h2 && yky_out && logdet_out, PSB_ERR_ARG, "NULL argument");
    PSB_REQUIRE(J > 0 && n_h2 > 0 && n_h2 <= 4096, PSB_ERR_ARG, "J = %d, n_h2 = %d out of range", J, n_h2);
    PSB_CUDA(cudaSetDevice(c->device));
    double *d = nullptr;
    const size_t nd = 2 * (size_t)J + 3 * (size_t)n_h2;
    PSB_CUDA(cudaMalloc(&d, nd * sizeof(double)));
    double *dS = d, *du = d + J, *dh = d + 2 * (size_t)J, *dout = dh + n_h2;
    cudaError_t e = cudaMemcpyAsync(dS, S, (size_t)J * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(du, uy2, (size_t)J * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dh, h2, (size_t)n_h2 * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    std::vector<double> out(2 * (size_t)n_h2);
    if (e == cudaSuccess) {
        k_nll_terms<<<n_h2, 256, 0, c->stream>>>(dS, du, J, dh, dout);
        c->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out.data(), dout, out.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    PSB_CUDA(e);
    for (int i = 0; i < n_h2; ++i) {
        yky_out[i] = out[2 * i];
        logdet_out[i] = out[2 * i + 1];
    }
    return PSB_OK;
}
