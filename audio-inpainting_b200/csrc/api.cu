// api.cu -- the C ABI of libainmf.so (include/ainmf.h): handle, tables, host RNG, workspace carving and the
// stage / whole-path entry points.  No torch types, no C++ exceptions across the boundary.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/ainmf.h"
#include "kernels.h"

using namespace ainmf;

namespace {

#ifdef AINMF_EMU
// ---- host RNG: numpy RandomState(seed).standard_normal --------------------------------------------------
// MT19937 seeded with init_genrand, 53-bit doubles, polar Box-Muller with the cached second deviate
// (numpy/random/src/legacy/legacy-distributions.c: legacy_gauss) -- what _initialize_nmf(init='random')
// draws from ($SP/sklearn/decomposition/_nmf.py:296-307).
struct Mt19937 {
    uint32_t mt[624];
    int pos;
    explicit Mt19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
    void refill() {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) refill();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double next_double() {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

void standard_normal_f32(uint32_t seed, float* out, size_t n) {
    Mt19937 g(seed);
    bool has = false;
    double cached = 0.0;
    for (size_t i = 0; i < n; ++i) {
        double v;
        if (has) {
            v = cached;
            has = false;
        } else {
            double x1, x2, r2;
            do {
                x1 = 2.0 * g.next_double() - 1.0;
                x2 = 2.0 * g.next_double() - 1.0;
                r2 = x1 * x1 + x2 * x2;
            } while (r2 >= 1.0 || r2 == 0.0);
            const double f = sqrt(-2.0 * log(r2) / r2);
            cached = f * x1;
            has = true;
            v = f * x2;
        }
        out[i] = (float)v;      // .astype(X.dtype)
    }
}

#endif  // AINMF_EMU (the CUDA build draws them on the device, rng.cu)

struct Tables {
    int n_fft = 0;
    float2* tw_half = nullptr;
    float2* tw_full = nullptr;
    float* window = nullptr;
    float win_sum = 0.f;
};

struct Normals {
    uint32_t seed = 0;
    int K = 0, T = 0, F = 0, t_begin = 0, t_count = 0;
    float* Hn = nullptr;   // [K][t_count] device: frames [t_begin, t_begin + t_count) of the (K, T) draw
    float* Wn = nullptr;   // [F][K] device
};

char g_create_error[256] = "";

}  // namespace

struct ainmf_context {
    int device = 0;
    int n_sm = 148;
    std::string err;
    std::vector<Tables> tables;
    std::vector<Normals> normals;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    int* poll_host = nullptr;     // pinned
    cudaEvent_t ev_poll[2] = {nullptr, nullptr};   // stop-flag polls, one group of iterations behind the launches
    // last chunk of ainmf_inpaint_host: ainmf_inpaint runs the inverse STFT in four parts and sends each part's waveform to
    // sink_y_host on sink_st as soon as it exists (the copy-out of the last chunk is the exposed one)
    float* sink_y_host = nullptr;
    cudaStream_t sink_st = nullptr;
    cudaEvent_t ev_part[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st_aux = nullptr;                 // side stream of the iteration (hbad next to the X.Ht kernel) and its fork / join events
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // ainmf_inpaint_host pipeline: copy-in / compute / copy-out streams and the events that order two chunks in flight
    cudaStream_t st_in = nullptr, st_cmp = nullptr, st_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // communicator for the time-frame-sharded mode (NCCL loaded at run time)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
};

namespace {

int fail(ainmf_handle h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else snprintf(g_create_error, sizeof g_create_error, "%s", buf);
    return code;
}
#define CU(h, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail(h, AINMF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
size_t al256(size_t n) { return (n + 255) / 256 * 256; }

int check_fft(ainmf_handle h, int n_fft, int hop, bool forward = true, bool inverse = true) {
    if (!is_pow2(n_fft) || n_fft < 64 || n_fft > 4096) return fail(h, AINMF_ERR_INVALID, "n_fft must be a power of two in [64, 4096], got %d", n_fft);
    if (hop <= 0 || hop % 4 != 0 || n_fft % hop != 0 || n_fft / hop > 8 || n_fft / hop < 1)
        return fail(h, AINMF_ERR_INVALID, "hop must be a multiple of 4 that divides n_fft with n_fft/hop <= 8, got %d", hop);
    if (stft_smem_need(n_fft, hop, forward, inverse) > 227 * 1024)
        return fail(h, AINMF_ERR_INVALID, "n_fft %d with hop %d needs %zu bytes of shared memory per block (limit 232448): use a smaller hop",
                    n_fft, hop, stft_smem_need(n_fft, hop, forward, inverse));
    return 0;
}

int geometry(long long N, int n_fft, int hop, StftGeom* g) {
    // scipy: x_ext = zeros(n/2) ++ x ++ zeros(n/2); nadd = (-(len(x_ext) - n) % hop) % n; T = (len + nadd - n)/hop + 1
    const long long ext = N + 2LL * (n_fft / 2);
    long long r = (ext - n_fft) % hop;
    long long nadd = ((hop - r) % hop) % n_fft;
    const long long T = (ext + nadd - n_fft) / hop + 1;
    if (T > 0x7fffffff / 2) return -1;
    g->N = N; g->n_fft = n_fft; g->hop = hop;
    g->T = (int)T; g->F = n_fft / 2 + 1; g->ldf = round_up(g->F, 4);
    return 0;
}

int get_tables(ainmf_handle h, int n_fft, FftTables* out) {
    for (const Tables& t : h->tables)
        if (t.n_fft == n_fft) { out->n_fft = n_fft; out->tw_half = t.tw_half; out->tw_full = t.tw_full; out->window = t.window; out->win_sum = t.win_sum; return 0; }
    const int M = n_fft / 2;
    std::vector<float2> th(M), tf(M + 1);
    std::vector<float> w(n_fft);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < M; ++k) { th[k].x = (float)cos(-2.0 * pi * k / M); th[k].y = (float)sin(-2.0 * pi * k / M); }
    for (int k = 0; k <= M; ++k) { tf[k].x = (float)cos(-2.0 * pi * k / n_fft); tf[k].y = (float)sin(-2.0 * pi * k / n_fft); }
    // get_window('hann_periodic') -> general_cosine(n, [0.5, 0.5], sym=False): fac = linspace(-pi, pi, n+1)[:n];
    // w = 0.5 + 0.5 cos(fac) in float64 ($SP/scipy/signal/windows/_windows.py:56-66), then cast to float32.
    float wsum = 0.f;
    {
        // float32 pairwise-free sum: numpy's pairwise sum of these values is exact (= n/2) for power-of-two n; we
        // accumulate in double and round once, which gives the same float.
        double acc = 0.0;
        for (int j = 0; j < n_fft; ++j) {
            const double fac = -pi + (2.0 * pi) * (double)j / (double)n_fft;
            w[j] = (float)(0.5 + 0.5 * cos(fac));
            acc += (double)w[j];
        }
        wsum = (float)acc;
    }
    Tables t;
    t.n_fft = n_fft; t.win_sum = wsum;
    CU(h, cudaMalloc((void**)&t.tw_half, sizeof(float2) * M));
    CU(h, cudaMalloc((void**)&t.tw_full, sizeof(float2) * (M + 1)));
    CU(h, cudaMalloc((void**)&t.window, sizeof(float) * n_fft));
    CU(h, cudaMemcpy(t.tw_half, th.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(t.tw_full, tf.data(), sizeof(float2) * (M + 1), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(t.window, w.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    h->tables.push_back(t);
    out->n_fft = n_fft; out->tw_half = t.tw_half; out->tw_full = t.tw_full; out->window = t.window; out->win_sum = wsum;
    return 0;
}

// The normals sklearn would draw for (seed, K, T, F): H (K,T) first, then W (F,K); of H only frames [t_begin, t_begin +
// t_count) are kept (a rank of the time-sharded mode needs its own frames only).  Generated on the device (rng.cu) on the
// caller's stream and cached in the handle; the host generator above serves the emulator build.
int get_normals(ainmf_handle h, uint32_t seed, int K, int T, int F, int t_begin, int t_count, const float** Wn, const float** Hn,
                cudaStream_t s) {
    for (const Normals& n : h->normals)
        if (n.seed == seed && n.K == K && n.T == T && n.F == F && n.t_begin == t_begin && n.t_count == t_count) { *Wn = n.Wn; *Hn = n.Hn; return 0; }
    if (h->normals.size() >= 4) {      // small cache: evict the oldest (after everything queued on it has run)
        cudaDeviceSynchronize();
        cudaFree(h->normals[0].Hn);
        cudaFree(h->normals[0].Wn);
        h->normals.erase(h->normals.begin());
    }
    const size_t nH = (size_t)K * t_count, nW = (size_t)F * K;
    Normals n;
    n.seed = seed; n.K = K; n.T = T; n.F = F; n.t_begin = t_begin; n.t_count = t_count;
    CU(h, cudaMalloc((void**)&n.Hn, sizeof(float) * nH));
    CU(h, cudaMalloc((void**)&n.Wn, sizeof(float) * nW));
#ifdef AINMF_EMU
    {
        std::vector<float> z((size_t)K * T + nW), hl(nH);
        standard_normal_f32(seed, z.data(), z.size());
        for (int k = 0; k < K; ++k) memcpy(hl.data() + (size_t)k * t_count, z.data() + (size_t)k * T + t_begin, sizeof(float) * t_count);
        CU(h, cudaMemcpy(n.Hn, hl.data(), sizeof(float) * nH, cudaMemcpyHostToDevice));
        CU(h, cudaMemcpy(n.Wn, z.data() + (size_t)K * T, sizeof(float) * nW, cudaMemcpyHostToDevice));
    }
#else
    CU(h, launch_numpy_normals(seed, K, T, t_begin, t_count, n.Hn, F, n.Wn, s));
#endif
    h->normals.push_back(n);
    *Wn = n.Wn; *Hn = n.Hn;
    return 0;
}

int get_scratch(ainmf_handle h, size_t bytes, void** out) {
    if (bytes > h->scratch_bytes) {
        if (h->scratch) cudaFree(h->scratch);
        h->scratch = nullptr; h->scratch_bytes = 0;
        CU(h, cudaMalloc(&h->scratch, bytes));
        h->scratch_bytes = bytes;
    }
    *out = h->scratch;
    return 0;
}

int get_pinned(ainmf_handle h, size_t bytes, void** out) {
    if (bytes > h->pinned_bytes) {
        if (h->pinned) cudaFreeHost(h->pinned);
        h->pinned = nullptr; h->pinned_bytes = 0;
        CU(h, cudaMallocHost(&h->pinned, bytes));
        h->pinned_bytes = bytes;
    }
    *out = h->pinned;
    return 0;
}

// ---- workspace of the whole path ---------------------------------------------------------------------------
struct Plan {
    StftGeom g;
    int B = 0, K = 0, KP = 0;
    ImputeWork iw;
    NmfWork nw;
    size_t off_V = 0, off_Z = 0, off_bad = 0, off_excl = 0, off_idx = 0, off_nbad = 0, off_nexcl = 0, off_fill = 0,
           off_state = 0, off_W = 0, off_Ht = 0, off_imp = 0, off_nmf = 0, off_notdone = 0, total = 0;
    bool compact = false;           // fit on good-first permuted frames (tensor-core path, n_outer == 1)
    size_t off_Xp = 0, off_Htp = 0, off_perm = 0, off_gidx = 0, off_gflags = 0, off_tgood = 0;
    long long vz_stride = 0, bad_stride = 0, w_stride = 0, h_stride = 0;
};

int make_plan(ainmf_handle h, const ainmf_params* p, Plan* pl) {
    if (!p) return fail(h, AINMF_ERR_INVALID, "params is NULL");
    if (p->batch <= 0) return fail(h, AINMF_ERR_INVALID, "batch must be positive, got %d", p->batch);
    int rc = check_fft(h, p->n_fft, p->hop);
    if (rc) return rc;
    if (p->n_samples < p->n_fft) return fail(h, AINMF_ERR_INVALID, "n_samples (%lld) must be >= n_fft (%d): scipy would shrink nperseg", (long long)p->n_samples, p->n_fft);
    if (p->rank < 1 || p->rank > 128) return fail(h, AINMF_ERR_INVALID, "rank must be in [1,128], got %d", p->rank);
    if (p->max_iter < 1) return fail(h, AINMF_ERR_INVALID, "max_iter must be >= 1");
    if (!(p->tol >= 0.f)) return fail(h, AINMF_ERR_INVALID, "tol must be >= 0");
    if (p->solver != AINMF_SOLVER_CD && p->solver != AINMF_SOLVER_MU && p->solver != AINMF_SOLVER_MU_KL) return fail(h, AINMF_ERR_INVALID, "unknown solver %d (0 = cd, 1 = mu, 2 = mu with the Kullback-Leibler divergence)", p->solver);
    if (p->n_outer < 1) return fail(h, AINMF_ERR_INVALID, "n_outer must be >= 1");
    if (geometry(p->n_samples, p->n_fft, p->hop, &pl->g)) return fail(h, AINMF_ERR_INVALID, "signal too long");
    if (p->col_start >= 0) {
        if (p->col_start < 1 || p->col_end <= p->col_start || p->col_end > pl->g.T)
            return fail(h, AINMF_ERR_INVALID, "need 1 <= col_start < col_end <= T (%d), got [%d,%d)", pl->g.T, p->col_start, p->col_end);
    } else {
        if (p->frac_den <= 0 || p->frac_num < 0 || p->frac_num >= p->frac_den) return fail(h, AINMF_ERR_INVALID, "need 0 <= frac_num < frac_den");
        if (!(p->threshold > 0.f)) return fail(h, AINMF_ERR_INVALID, "threshold must be positive");
    }
    const int B = p->batch, T = pl->g.T, F = pl->g.F, ldf = pl->g.ldf;
    pl->B = B; pl->K = p->rank; pl->KP = ainmf_padded_rank(p->rank);
    const int KP = pl->KP;
    impute_plan(T, &pl->iw);
    nmf_plan(B, T, F, KP, h->n_sm, &pl->nw);
    if (p->solver != AINMF_SOLVER_CD) { pl->nw.want_mu = (p->solver == AINMF_SOLVER_MU_KL) ? 2 : 1; pl->nw.use_tc = 0; pl->nw.exact_viol = 0; }
    pl->vz_stride = (long long)T * ldf;
    pl->bad_stride = round_up(T, 16);
    pl->w_stride = (long long)F * KP;
    pl->h_stride = (long long)T * KP;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
    pl->off_V = take(sizeof(float) * (size_t)B * pl->vz_stride);
    pl->off_Z = take(sizeof(float2) * (size_t)B * pl->vz_stride);
    pl->off_bad = take((size_t)B * pl->bad_stride);
    pl->off_excl = take((size_t)B * pl->bad_stride);
    pl->off_idx = take(sizeof(int) * (size_t)B * T);
    pl->off_nbad = take(sizeof(int) * (size_t)B);
    pl->off_nexcl = take(sizeof(int) * (size_t)B);
    pl->off_fill = take(sizeof(float) * (size_t)B * ldf);
    pl->off_state = take(sizeof(ClipState) * (size_t)B);
    pl->off_W = take(sizeof(float) * (size_t)B * pl->w_stride);
    pl->off_Ht = take(sizeof(float) * (size_t)B * pl->h_stride);
    pl->off_imp = take(impute_work_bytes(B, F, pl->iw));
    pl->off_nmf = take(nmf_work_bytes(B, T, F, KP, pl->nw));
    pl->off_notdone = take(sizeof(int) * 4);
    {
        const char* nc = getenv("AINMF_NO_COMPACT");
        // worth it once the tiles fill the machine (a single short clip only pays the extra small launches)
        pl->compact = pl->nw.use_tc && p->n_outer == 1 && p->solver == AINMF_SOLVER_CD && !(nc && nc[0] == '1') &&
                      (long long)B * ceil_div(T, 128) >= h->n_sm;
    }
    if (pl->compact) {
        pl->off_Xp = take(sizeof(float) * (size_t)B * pl->vz_stride);
        pl->off_Htp = take(sizeof(float) * (size_t)B * pl->h_stride);
        pl->off_perm = take(sizeof(int) * (size_t)B * T);
        pl->off_gidx = take(sizeof(int) * (size_t)B * T);
        pl->off_gflags = take((size_t)B * pl->bad_stride);
        pl->off_tgood = take(sizeof(int) * (size_t)B);
    }
    pl->total = o;
    return 0;
}

__global__ void count_not_done_kernel(const ClipState* st, int B, int* out) {
    // single block
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) c += st[b].done ? 0 : 1;
    if (c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) out[0] = s_cnt;
}
__global__ void status_summary_kernel(const ClipState* st, int B, int* out) {
    // out[1] = #clips with status 2 (all frames bad), out[2] = #clips with status 0 (work to do)
    __shared__ int s_a, s_b;
    if (threadIdx.x == 0) { s_a = 0; s_b = 0; }
    __syncthreads();
    int a = 0, w = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) { a += st[b].status == 2; w += st[b].status == 0; }
    if (a) atomicAdd(&s_a, a);
    if (w) atomicAdd(&s_b, w);
    __syncthreads();
    if (threadIdx.x == 0) { out[1] = s_a; out[2] = s_b; }
}

// Clips whose every frame is flagged (status 2: the fill spectrum is the mean of an empty set, NaN in the reference) are
// passed through: the inverse transform copies x when its frame count reads 0.  The exported n_bad stays T.
__global__ void pass_through_all_bad_kernel(const ClipState* st, int B, int* n_bad_for_istft) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && st[b].status == 2) n_bad_for_istft[b] = 0;
}

__global__ void copy_indices_kernel(const int* __restrict__ idx, const ClipState* __restrict__ st, int T, int* __restrict__ out) {
    const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) out[(long long)b * T + t] = (t < st[b].n_bad) ? idx[(long long)b * T + t] : -1;
}

// Runs up to max_iter iterations.  When tol > 0 the number of clips still iterating is read back every `poll` iterations,
// but the host looks at the answer of group g only after it has queued group g + 1: the stream never drains while the host
// waits and launches.  Kernels skip clips whose stop rule fired (ClipState.done, set on the device), so the up to `poll`
// extra iterations queued after the last clip converged do nothing and every clip's n_iter is exact.
// Side stream of the coordinate-descent iteration (tensor-core path, good-first frame order): hbad_kernel and its reduce
// read Ht only, as the X.Ht kernel does, and run next to it (iterate_impl forks and joins with the two events).
int attach_aux_stream(ainmf_handle h, const NmfProblem& prob, NmfWork* nw) {
    if (!nw->use_tc || !prob.t_good || getenv("AINMF_NO_AUX_STREAM")) return 0;
    if (!h->st_aux) {
        CU(h, cudaStreamCreateWithFlags(&h->st_aux, cudaStreamNonBlocking));
        CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    nw->aux_stream = h->st_aux; nw->ev_fork = h->ev_fork; nw->ev_join = h->ev_join;
    return 0;
}

int run_iterations(ainmf_handle h, const NmfProblem& prob, const NmfWork& nw, int max_iter, float tol, int* d_flag,
                   cudaStream_t s) {
    const bool kl = prob.solver == AINMF_SOLVER_MU_KL;
    const bool mu = prob.solver == AINMF_SOLVER_MU || kl;
    const int poll = mu ? 10 : 8;          // MU tests convergence every 10th iteration only
    if (!mu && nmf_coop_eligible(prob, nw, h->n_sm) && !getenv("AINMF_NO_COOP")) {
        // one clip: the whole fit is one cooperative launch, stop rule on the device, no polls (nmf_coop.cu)
        prof_begin(PROF_H_STEP, s);                       // the profile's H-step slot times the whole fit (one launch)
        CU(h, nmf_coop_fit(prob, nw, max_iter, h->n_sm, s));
        prof_end(PROF_H_STEP, s);
        return 0;
    }
    if (mu && tol > 0.f) CU(h, kl ? nmf_mukl_begin(prob, nw, s) : nmf_mu_begin(prob, nw, s));
    if (tol > 0.f && !h->ev_poll[0])
        for (int i = 0; i < 2; ++i) CU(h, cudaEventCreateWithFlags(&h->ev_poll[i], cudaEventDisableTiming));
    NmfWork nwf = nw;                      // with the side stream for the reductions that run next to the X.Ht kernel
    if (!mu && attach_aux_stream(h, prob, &nwf)) return AINMF_ERR_CUDA;
    int pending = -1;
    for (int it = 1; it <= max_iter; ++it) {
        if (kl) CU(h, nmf_mukl_iterate(prob, nw, it, s));
        else if (mu) CU(h, nmf_mu_iterate(prob, nw, it, s));
        else CU(h, nmf_cd_iterate(prob, nwf, it, s));
        if (tol > 0.f && (it % poll == 0) && it < max_iter) {
            const int g = (it / poll) & 1;
            int* d_out = d_flag + (g ? 3 : 0);
            AINMF_LAUNCH(count_not_done_kernel, dim3(1), dim3(kThreads), 0, s, prob.state, prob.B, d_out);
            CU(h, cudaGetLastError());
            CU(h, cudaMemcpyAsync(h->poll_host + 8 + g, d_out, sizeof(int), cudaMemcpyDeviceToHost, s));
            CU(h, cudaEventRecord(h->ev_poll[g], s));
            if (pending >= 0) {
                CU(h, cudaEventSynchronize(h->ev_poll[pending]));
                if (h->poll_host[8 + pending] == 0) break;
            }
            pending = g;
        }
    }
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* ainmf_version(void) { return "ainmf 0.1 (sm_100a)"; }

void ainmf_params_default(ainmf_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->batch = 1; p->n_samples = 0; p->n_fft = 1024; p->hop = 256; p->rank = 40; p->max_iter = 200; p->tol = 1e-4f;
    p->solver = AINMF_SOLVER_CD; p->seed = 42; p->threshold = 1e-4f; p->frac_num = 9; p->frac_den = 10;
    p->col_start = -1; p->col_end = -1; p->n_outer = 1;
}

int32_t ainmf_padded_rank(int32_t rank) { return rank <= 32 ? 32 : (rank <= 64 ? 64 : 128); }

int ainmf_create(ainmf_handle* out, int device) {
    if (!out) return fail(nullptr, AINMF_ERR_INVALID, "out is NULL");
    *out = nullptr;
#ifndef AINMF_EMU
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) return fail(nullptr, AINMF_ERR_NO_DEVICE, "no CUDA device (%s); ainmf has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, AINMF_ERR_INVALID, "device %d out of range (%d devices)", device, count);
#endif
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(nullptr, AINMF_ERR_CUDA, "cannot select device %d", device);
#ifndef AINMF_EMU
    if (prop.major != 10) return fail(nullptr, AINMF_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
#endif
    ainmf_context* h = new (std::nothrow) ainmf_context();
    if (!h) return fail(nullptr, AINMF_ERR_INVALID, "out of host memory");
    h->device = device;
    h->n_sm = prop.multiProcessorCount;
    if (cudaMallocHost((void**)&h->poll_host, 64) != cudaSuccess) { delete h; return fail(nullptr, AINMF_ERR_CUDA, "cudaMallocHost failed"); }
    *out = h;
    return AINMF_OK;
}

int ainmf_comm_destroy_internal(ainmf_handle h);

int ainmf_destroy(ainmf_handle h) {
    if (!h) return AINMF_OK;
    cudaSetDevice(h->device);
    ainmf_comm_destroy_internal(h);
    for (Tables& t : h->tables) { cudaFree(t.tw_half); cudaFree(t.tw_full); cudaFree(t.window); }
    for (Normals& n : h->normals) { cudaFree(n.Hn); cudaFree(n.Wn); }
    if (h->scratch) cudaFree(h->scratch);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_poll[i]) cudaEventDestroy(h->ev_poll[i]);
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->st_in) cudaStreamDestroy(h->st_in);
    if (h->st_cmp) cudaStreamDestroy(h->st_cmp);
    if (h->st_out) cudaStreamDestroy(h->st_out);
    for (int i = 0; i < 4; ++i) if (h->ev_part[i]) cudaEventDestroy(h->ev_part[i]);
    if (h->st_aux) { cudaStreamDestroy(h->st_aux); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->poll_host) cudaFreeHost(h->poll_host);
    delete h;
    return AINMF_OK;
}

const char* ainmf_last_error(ainmf_handle h) { return h ? h->err.c_str() : g_create_error; }

int ainmf_set_window(ainmf_handle h, int32_t n_fft, const float* window_host) {
    if (!h) return AINMF_ERR_INVALID;
    if (!is_pow2(n_fft) || n_fft < 64 || n_fft > 4096) return fail(h, AINMF_ERR_INVALID, "n_fft must be a power of two in [64, 4096], got %d", n_fft);
    CU(h, cudaSetDevice(h->device));
    FftTables tb;
    int rc = get_tables(h, n_fft, &tb);                   // creates the entry (periodic Hann) on first use
    if (rc) return rc;
    for (Tables& t : h->tables) {
        if (t.n_fft != n_fft) continue;
        std::vector<float> w(n_fft);
        double acc = 0.0;
        const double pi = 3.14159265358979323846;
        for (int j = 0; j < n_fft; ++j) {
            w[j] = window_host ? window_host[j] : (float)(0.5 + 0.5 * cos(-pi + (2.0 * pi) * (double)j / (double)n_fft));
            if (!(w[j] == w[j]) || fabsf(w[j]) > 3.0e38f) return fail(h, AINMF_ERR_INVALID, "window[%d] is not finite", j);
            acc += (double)w[j];
        }
        if (!(fabs(acc) > 0.0)) return fail(h, AINMF_ERR_INVALID, "the window sums to zero: scipy's 1/sum(window) scaling is undefined");
        CU(h, cudaDeviceSynchronize());
        CU(h, cudaMemcpy(t.window, w.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
        t.win_sum = (float)acc;                           // float32 sum of the float32 window ($SP/scipy/signal/_spectral_py.py:2272-2282)
    }
    return AINMF_OK;
}

int ainmf_standard_normal(ainmf_handle h, uint32_t seed, int64_t n, float* out, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!out || n < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_standard_normal");
    CU(h, cudaSetDevice(h->device));
#ifdef AINMF_EMU
    std::vector<float> z((size_t)n);
    standard_normal_f32(seed, z.data(), z.size());
    CU(h, cudaMemcpy(out, z.data(), sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
#else
    CU(h, launch_numpy_normals(seed, 1, n, 0, n, out, 0, nullptr, (cudaStream_t)stream));
#endif
    return AINMF_OK;
}

int ainmf_stft_geometry(int64_t n_samples, int32_t n_fft, int32_t hop, int32_t* T, int32_t* F, int32_t* ldf) {
    if (!is_pow2(n_fft) || hop <= 0 || n_fft % hop != 0 || n_samples < 1) return AINMF_ERR_INVALID;
    StftGeom g;
    if (geometry(n_samples, n_fft, hop, &g)) return AINMF_ERR_INVALID;
    if (T) *T = g.T;
    if (F) *F = g.F;
    if (ldf) *ldf = g.ldf;
    return AINMF_OK;
}

int ainmf_stft(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, int32_t n_fft, int32_t hop,
               float* mag_ft, float* Z_ft, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || batch <= 0) return fail(h, AINMF_ERR_INVALID, "x is NULL or batch <= 0");
    int rc = check_fft(h, n_fft, hop, true, false);
    if (rc) return rc;
    if (n_samples < n_fft) return fail(h, AINMF_ERR_INVALID, "n_samples (%lld) must be >= n_fft (%d)", (long long)n_samples, n_fft);
    StftGeom g;
    if (geometry(n_samples, n_fft, hop, &g)) return fail(h, AINMF_ERR_INVALID, "signal too long");
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    FftTables tb;
    if ((rc = get_tables(h, n_fft, &tb))) return rc;
    const long long vz = (long long)g.T * g.ldf;
    void* scr;
    if ((rc = get_scratch(h, al256(sizeof(float) * batch * vz) + sizeof(float2) * batch * vz, &scr))) return rc;
    float* V = (float*)scr;
    float2* Z = (float2*)((char*)scr + al256(sizeof(float) * batch * vz));
    CU(h, launch_stft(x, n_samples, 0, n_samples, batch, g, 0, g.T, tb, V, Z, vz, s));
    if (mag_ft) CU(h, launch_transpose_f32(V, vz, g.ldf, g.T, g.F, mag_ft, (long long)g.F * g.T, g.T, 0, batch, s));
    if (Z_ft) CU(h, launch_transpose_c64(Z, vz, g.ldf, g.T, g.F, (float2*)Z_ft, (long long)g.F * g.T, g.T, 0, batch, s));
    return AINMF_OK;
}

int ainmf_gap_mask(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, int32_t hop, int32_t n_frames,
                   float threshold, int32_t frac_num, int32_t frac_den, uint8_t* bad, int32_t* bad_idx,
                   int32_t* n_bad, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || batch <= 0 || n_samples < 1 || hop < 1 || n_frames < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_gap_mask");
    if (frac_den <= 0 || frac_num < 0) return fail(h, AINMF_ERR_INVALID, "need frac_den > 0 and frac_num >= 0");
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    unsigned char* flags = bad;
    if (!flags) {
        void* scr;
        int rc = get_scratch(h, (size_t)batch * n_frames, &scr);
        if (rc) return rc;
        flags = (unsigned char*)scr;
    }
    CU(h, launch_gap_mask(x, n_samples, 0, n_samples, batch, n_samples, hop, 0, n_frames, threshold, frac_num, frac_den,
                          flags, n_frames, s));
    if (bad_idx && n_bad) CU(h, launch_compact(flags, n_frames, batch, n_frames, bad_idx, n_frames, n_bad, s));
    else if (bad_idx || n_bad) return fail(h, AINMF_ERR_INVALID, "bad_idx and n_bad must be given together");
    return AINMF_OK;
}

int ainmf_nmf_fit(ainmf_handle h, const float* X_ft, int32_t batch, int32_t F, int32_t T, int32_t rank,
                  int32_t max_iter, float tol, int32_t solver, uint32_t seed, const float* W0, const float* H0,
                  float* W, float* H, float* err, int32_t* n_iter, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!X_ft || batch <= 0 || F < 1 || T < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_nmf_fit");
    if (rank < 1 || rank > 128) return fail(h, AINMF_ERR_INVALID, "rank must be in [1,128], got %d", rank);
    if (max_iter < 1 || !(tol >= 0.f)) return fail(h, AINMF_ERR_INVALID, "need max_iter >= 1 and tol >= 0");
    if (solver != AINMF_SOLVER_CD && solver != AINMF_SOLVER_MU && solver != AINMF_SOLVER_MU_KL) return fail(h, AINMF_ERR_INVALID, "unknown solver %d (0 = cd, 1 = mu, 2 = mu with the Kullback-Leibler divergence)", solver);
    if ((W0 == nullptr) != (H0 == nullptr)) return fail(h, AINMF_ERR_INVALID, "W0 and H0 must be given together");
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    const int KP = ainmf_padded_rank(rank), ldf = round_up(F, 4), B = batch;
    ImputeWork iw;
    NmfWork nw;
    impute_plan(T, &iw);
    nmf_plan(B, T, F, KP, h->n_sm, &nw);
    if (solver != AINMF_SOLVER_CD) { nw.want_mu = (solver == AINMF_SOLVER_MU_KL) ? 2 : 1; nw.use_tc = 0; nw.exact_viol = 0; }
    const long long xs = (long long)T * ldf, ws = (long long)F * KP, hs = (long long)T * KP;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
    const size_t oX = take(sizeof(float) * B * xs), oW = take(sizeof(float) * B * ws), oH = take(sizeof(float) * B * hs);
    const size_t oS = take(sizeof(ClipState) * B), oI = take(impute_work_bytes(B, F, iw)), oN = take(nmf_work_bytes(B, T, F, KP, nw));
    const size_t oB = take((size_t)B * round_up(T, 16)), oF = take(64);
    void* scr;
    int rc = get_scratch(h, o, &scr);
    if (rc) return rc;
    char* base = (char*)scr;
    NmfProblem pr;
    pr.B = B; pr.T = T; pr.F = F; pr.ldf = ldf; pr.KP = KP; pr.tol = tol; pr.solver = solver;
    pr.Xt = (float*)(base + oX); pr.x_stride = xs;
    pr.W = (float*)(base + oW); pr.w_stride = ws;
    pr.Ht = (float*)(base + oH); pr.h_stride = hs;
    pr.state = (ClipState*)(base + oS);
    impute_carve(base + oI, B, F, &iw);
    nmf_carve(base + oN, B, T, F, KP, &nw);
    unsigned char* nobad = (unsigned char*)(base + oB);
    TcMaps tcm;
    if (nmf_tc_setup(pr, &nw, &tcm)) return fail(h, AINMF_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    if (nw.zero_flags) CU(h, cudaMemsetAsync(nw.zero_flags, 0, (size_t)B * nw.zero_stride, s));
    CU(h, cudaMemsetAsync(nw.counters, 0, sizeof(unsigned) * B, s));
    CU(h, cudaMemsetAsync(nobad, 0, (size_t)B * round_up(T, 16), s));
    CU(h, cudaMemsetAsync(pr.state, 0, sizeof(ClipState) * B, s));
    // (F,T) -> [T][ldf], pad bins zeroed
    CU(h, launch_transpose_f32(X_ft, (long long)F * T, T, F, T, pr.Xt, xs, ldf, 1, B, s));
    CU(h, launch_colsums(pr.Xt, xs, ldf, F, T, B, nullptr, 0, iw, s));
    CU(h, launch_mean(F, T, B, pr.state, iw, s));
    if (W0) {
        CU(h, launch_pack_factors(W0, H0, B, F, T, rank, KP, pr.W, ws, pr.Ht, hs, s));
    } else {
        const float *Wn, *Hn;
        if ((rc = get_normals(h, seed, rank, T, F, 0, T, &Wn, &Hn, s))) return rc;
        CU(h, launch_init_factors(Wn, Hn, T, 0, B, F, T, rank, KP, pr.state, pr.W, ws, pr.Ht, hs, s));
    }
    if ((rc = run_iterations(h, pr, nw, max_iter, tol, (int*)(base + oF), s))) return rc;
    if (solver == AINMF_SOLVER_MU_KL) CU(h, nmf_mukl_error(pr, nw, true, s));      // reconstruction_err_ = sqrt(2 D_KL)
    CU(h, nmf_finalize(pr, nw, nobad, round_up(T, 16), s));
    if (solver == AINMF_SOLVER_MU_KL) CU(h, nmf_mukl_set_err(pr, nw, s));
    CU(h, launch_unpack_factors(pr.W, ws, pr.Ht, hs, B, F, T, rank, KP, W, H, s));
    CU(h, launch_export_state(pr.state, B, nullptr, n_iter, err, nullptr, s));
    return AINMF_OK;
}

int ainmf_istft(ainmf_handle h, const float* Z_ft, int32_t batch, int32_t T, int32_t n_fft, int32_t hop,
                int64_t n_samples, float* y, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!Z_ft || !y || batch <= 0) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_istft");
    int rc = check_fft(h, n_fft, hop, false, true);
    if (rc) return rc;
    StftGeom g;
    g.N = n_samples; g.n_fft = n_fft; g.hop = hop; g.T = T; g.F = n_fft / 2 + 1; g.ldf = round_up(g.F, 4);
    if (T < 1 || n_samples < 1 || n_samples > (long long)(T - 1) * hop)
        return fail(h, AINMF_ERR_INVALID, "n_samples must be in [1, (T-1)*hop]");
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    FftTables tb;
    if ((rc = get_tables(h, n_fft, &tb))) return rc;
    const long long vz = (long long)T * g.ldf;
    const size_t bs = round_up(T, 16);
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
    const size_t oZ = take(sizeof(float2) * batch * vz), oB = take((size_t)batch * bs), oN = take(sizeof(int) * batch);
    void* scr;
    if ((rc = get_scratch(h, o, &scr))) return rc;
    char* base = (char*)scr;
    float2* Z = (float2*)(base + oZ);
    unsigned char* bad = (unsigned char*)(base + oB);
    int* nb = (int*)(base + oN);
    CU(h, cudaMemsetAsync(bad, 0, (size_t)batch * bs, s));
    CU(h, cudaMemsetAsync(nb, 1, sizeof(int) * batch, s));     // non-zero: take the transform path
    CU(h, launch_transpose_c64((const float2*)Z_ft, (long long)g.F * T, T, g.F, T, Z, vz, g.ldf, 1, batch, s));
    CU(h, launch_istft(nullptr, Z, vz, bad, bs, nb, nullptr, 0, 0, batch, g, 0, T, tb, y, n_samples, 0, n_samples, T, s));
    return AINMF_OK;
}

size_t ainmf_workspace_bytes(ainmf_handle h, const ainmf_params* p) {
    if (!h) return 0;
    Plan pl;
    if (make_plan(h, p, &pl)) return 0;
    return pl.total;
}

int ainmf_inpaint(ainmf_handle h, const ainmf_params* p, const float* x, const float* W0, const float* H0, float* y,
                  int32_t* bad_idx, int32_t* n_bad, float* W, float* H, float* err, int32_t* n_iter,
                  void* workspace, size_t workspace_bytes, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    Plan pl;
    int rc = make_plan(h, p, &pl);
    if (rc) return rc;
    if (!x || !y) return fail(h, AINMF_ERR_INVALID, "x and y must not be NULL");
    if ((W0 == nullptr) != (H0 == nullptr)) return fail(h, AINMF_ERR_INVALID, "W0 and H0 must be given together");
    if (!workspace || workspace_bytes < pl.total)
        return fail(h, AINMF_ERR_WORKSPACE, "workspace of %zu bytes needed, %zu given", pl.total, workspace_bytes);
    if (((uintptr_t)workspace & 255) != 0) return fail(h, AINMF_ERR_INVALID, "workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, cudaSetDevice(h->device));
    const StftGeom& g = pl.g;
    const int B = pl.B, T = g.T, F = g.F, ldf = g.ldf, K = pl.K, KP = pl.KP;
    const long long N = g.N;
    FftTables tb;
    if ((rc = get_tables(h, g.n_fft, &tb))) return rc;
    char* base = (char*)workspace;
    float* V = (float*)(base + pl.off_V);
    float2* Z = (float2*)(base + pl.off_Z);
    unsigned char* bad = (unsigned char*)(base + pl.off_bad);
    unsigned char* excl = (unsigned char*)(base + pl.off_excl);
    int* idx = (int*)(base + pl.off_idx);
    int* d_nbad = (int*)(base + pl.off_nbad);
    int* d_nexcl = (int*)(base + pl.off_nexcl);
    float* fill = (float*)(base + pl.off_fill);
    ClipState* st = (ClipState*)(base + pl.off_state);
    int* d_flag = (int*)(base + pl.off_notdone);
    impute_carve(base + pl.off_imp, B, F, &pl.iw);
    nmf_carve(base + pl.off_nmf, B, T, F, KP, &pl.nw);
    NmfProblem pr;
    pr.B = B; pr.T = T; pr.F = F; pr.ldf = ldf; pr.KP = KP; pr.tol = p->tol; pr.solver = p->solver;
    pr.Xt = V; pr.x_stride = pl.vz_stride;
    pr.W = (float*)(base + pl.off_W); pr.w_stride = pl.w_stride;
    pr.Ht = (float*)(base + pl.off_Ht); pr.h_stride = pl.h_stride;
    pr.state = st;
    // the fit itself may run on a good-first permutation of the frames (prf); everything else uses the original order
    NmfProblem prf = pr;
    int* perm = nullptr; int* d_tgood = nullptr;
    if (pl.compact) {
        prf.Xt = (float*)(base + pl.off_Xp);
        prf.Ht = (float*)(base + pl.off_Htp);
        perm = (int*)(base + pl.off_perm);
        d_tgood = (int*)(base + pl.off_tgood);
        prf.t_good = d_tgood; prf.fill = fill; prf.fill_stride = ldf;
    }
    TcMaps tcm;
    if (nmf_tc_setup(prf, &pl.nw, &tcm)) return fail(h, AINMF_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    if (pl.nw.zero_flags) CU(h, cudaMemsetAsync(pl.nw.zero_flags, 0, (size_t)B * pl.nw.zero_stride, s));

    CU(h, cudaMemsetAsync(pl.nw.counters, 0, sizeof(unsigned) * B, s));
    // a4: STFT
    CU(h, launch_stft(x, N, 0, N, B, g, 0, T, tb, V, Z, pl.vz_stride, s));
    // a2/a3: frame mask
    const unsigned char* excl_flags = bad;
    const int* d_nex = d_nbad;
    if (p->col_start >= 0) {
        CU(h, launch_range_mask(B, T, p->col_start, T, excl, pl.bad_stride, s));       // fill = mean of frames < col_start
        CU(h, launch_compact(excl, pl.bad_stride, B, T, idx, T, d_nexcl, s));
        CU(h, launch_range_mask(B, T, p->col_start, p->col_end, bad, pl.bad_stride, s));
        excl_flags = excl;
        d_nex = d_nexcl;
    } else {
        CU(h, launch_gap_mask(x, N, 0, N, B, N, g.hop, 0, T, p->threshold, p->frac_num, p->frac_den, bad, pl.bad_stride, s));
    }
    CU(h, launch_compact(bad, pl.bad_stride, B, T, idx, T, d_nbad, s));
    // a5: imputation (in place: V becomes X)
    CU(h, launch_colsums(V, pl.vz_stride, ldf, F, T, B, excl_flags, pl.bad_stride, pl.iw, s));
    CU(h, launch_fill(V, pl.vz_stride, ldf, F, T, T, B, bad, pl.bad_stride, d_nbad, d_nex, fill, st, pl.iw, s));
    // one early look at the counts: rejects the undefined all-bad case and skips the fit when nothing is bad
    AINMF_LAUNCH(status_summary_kernel, dim3(1), dim3(kThreads), 0, s, st, B, d_flag);
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(h->poll_host, d_flag, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    // a single clip with every frame flagged fails like the reference does for that file; inside a batch such a clip is
    // passed through (y = x, err = NaN, n_iter = 0, n_bad = T) and the other clips are restored
    if (h->poll_host[1] > 0 && B == 1) return fail(h, AINMF_ERR_ALL_BAD, "every frame is flagged: the fill spectrum is undefined");
    if (h->poll_host[1] > 0) {
        AINMF_LAUNCH(pass_through_all_bad_kernel, dim3(ceil_div(B, kThreads)), dim3(kThreads), 0, s, st, B, d_nbad);
        CU(h, cudaGetLastError());
    }
    const bool any_work = h->poll_host[2] > 0;
    if (any_work) {
        const float *Wn = nullptr, *Hn = nullptr;
        if (!W0 && (rc = get_normals(h, p->seed, K, T, F, 0, T, &Wn, &Hn, s))) return rc;
        // a small spectrogram (config 1: 257 x 19) runs all its refits in one launch with everything in shared memory
        const bool small = !W0 && p->solver == AINMF_SOLVER_CD && !pl.compact && !pl.nw.use_tc && nmf_small_eligible(F, T, K) &&
                           !getenv("AINMF_NO_SMALL");
        if (small) CU(h, nmf_small_refit(pr, K, Wn, Hn, bad, pl.bad_stride, p->n_outer, p->max_iter, s));
        for (int outer = 0; outer < (small ? 0 : p->n_outer); ++outer) {
            // a6: initial factors from mean(X) (sklearn draws a fresh RandomState(seed) at every fit)
            CU(h, launch_colsums(V, pl.vz_stride, ldf, F, T, B, nullptr, 0, pl.iw, s));
            CU(h, launch_mean(F, T, B, st, pl.iw, s));
            if (W0) CU(h, launch_pack_factors(W0, H0, B, F, T, K, KP, pr.W, pr.w_stride, pr.Ht, pr.h_stride, s));
            else CU(h, launch_init_factors(Wn, Hn, T, 0, B, F, T, K, KP, st, pr.W, pr.w_stride, pr.Ht, pr.h_stride, s));
            // a7: the fit
            if (pl.compact) {
                unsigned char* gflags = (unsigned char*)(base + pl.off_gflags);
                int* gidx = (int*)(base + pl.off_gidx);
                CU(h, launch_invert_flags(bad, pl.bad_stride, B, T, gflags, s));
                CU(h, launch_compact(gflags, pl.bad_stride, B, T, gidx, T, d_tgood, s));
                CU(h, launch_build_perm(gidx, d_tgood, idx, B, T, perm, s));
                CU(h, launch_gather_rows(V, pl.vz_stride, prf.Xt, pl.vz_stride, ldf, perm, B, T, d_tgood, s));
                CU(h, launch_gather_rows(pr.Ht, pr.h_stride, prf.Ht, pr.h_stride, KP, perm, B, T, nullptr, s));
            }
            if ((rc = run_iterations(h, prf, pl.nw, p->max_iter, p->tol, d_flag, s))) return rc;
            if (pl.compact) CU(h, launch_scatter_rows(prf.Ht, pr.h_stride, pr.Ht, pr.h_stride, KP, perm, B, T, s));
            // a8 + a9: objective, then bad frames <- (W H) frames
            if (p->solver == AINMF_SOLVER_MU_KL) CU(h, nmf_mukl_error(pr, pl.nw, true, s));
            CU(h, nmf_finalize(pr, pl.nw, bad, pl.bad_stride, s));
            if (p->solver == AINMF_SOLVER_MU_KL) CU(h, nmf_mukl_set_err(pr, pl.nw, s));
        }
    }
    // a10 + a11: recombine with the corrupted phase, inverse STFT, trim
    if (h->sink_y_host && B >= 8) {
        for (int q = 0; q < 4; ++q) {
            const long long b0 = (long long)B * q / 4, b1 = (long long)B * (q + 1) / 4;
            CU(h, launch_istft(V + b0 * pl.vz_stride, Z + b0 * pl.vz_stride, pl.vz_stride, bad + b0 * pl.bad_stride, pl.bad_stride,
                               d_nbad + b0, x + b0 * N, N, 0, (int)(b1 - b0), g, 0, T, tb, y + b0 * N, N, 0, N, T, s));
            CU(h, cudaEventRecord(h->ev_part[q], s));
            CU(h, cudaStreamWaitEvent(h->sink_st, h->ev_part[q], 0));
            CU(h, cudaMemcpyAsync(h->sink_y_host + b0 * N, y + b0 * N, sizeof(float) * (size_t)(b1 - b0) * N, cudaMemcpyDeviceToHost, h->sink_st));
        }
    } else {
        CU(h, launch_istft(V, Z, pl.vz_stride, bad, pl.bad_stride, d_nbad, x, N, 0, B, g, 0, T, tb, y, N, 0, N, T, s));
    }
    if (bad_idx) {          // entries at and beyond n_bad are -1, whatever the workspace held
        AINMF_LAUNCH(copy_indices_kernel, dim3(ceil_div(T, kThreads), B), dim3(kThreads), 0, s, idx, st, T, bad_idx);
        CU(h, cudaGetLastError());
    }
    CU(h, launch_export_state(st, B, n_bad, n_iter, err, nullptr, s));
    if (any_work) CU(h, launch_unpack_factors(pr.W, pr.w_stride, pr.Ht, pr.h_stride, B, F, T, K, KP, W, H, s));
    return AINMF_OK;
}

}  // extern "C"

namespace {
// Chunk schedule of the host entry points (see inpaint_host_impl): c_begin = first clip of every chunk, then the batch size.
void plan_host_chunks(long long batch, long long cap, long long n_sm, std::vector<long long>* c_begin) {
    if (cap > 512) cap = 512;
    if (cap > batch) cap = batch;
    long long first = cap;
    if (batch >= 2 * n_sm && cap >= 2 * n_sm) first = n_sm;
    if (const char* ev = getenv("AINMF_HOST_CHUNK")) {      // development switch: clips per chunk, all equal
        const long long v = atoll(ev);
        if (v >= 1 && v < cap) first = cap = v;
    }
    if (const char* ev = getenv("AINMF_HOST_FIRST")) {      // development switch: clips in the first chunk
        const long long v = atoll(ev);
        if (v >= 1 && v <= cap) first = v;
    }
    c_begin->clear();
    c_begin->push_back(0);
    const long long done = first < batch ? first : batch;
    c_begin->push_back(done);
    // the kernels hand out per-clip items to n_sm (or 2 n_sm) CTAs: a chunk of m x n_sm clips fills every round, 444 clips
    // cost 0.574 ms each where 494 cost 0.591 and 512 0.587.  What is left at the end goes with the last chunk if the cap
    // allows (a tail of 68 clips runs at 0.9 ms per clip), else the last multiple is shortened to leave n_sm or more.
    const long long sm = n_sm, mult = cap >= sm ? cap / sm * sm : cap;
    for (long long at = done; at < batch;) {
        const long long left = batch - at;
        long long take_n = left;
        if (left > cap) {
            take_n = mult;
            if (left - take_n < sm && take_n > sm) take_n -= sm;
        }
        at += take_n;
        c_begin->push_back(at);
    }
}

// ainmf_inpaint_host (channels == 0: float32 in x_host / y_host) and ainmf_inpaint_host_pcm16 (channels >= 1: interleaved
// int16 in pcm_in_host, int16 in pcm_out_host; the chunk is converted on the device either side of the fit, pcm.cu)
int inpaint_host_impl(ainmf_handle h, const ainmf_params* p, const float* x_host, float* y_host, const int16_t* pcm_in_host,
                      int channels, int16_t* pcm_out_host, float* peak_host, int32_t* n_bad_host, float* err_host,
                      int32_t* n_iter_host, size_t max_device_bytes) {
    const bool pcm = channels > 0;
    const auto t_call = std::chrono::steady_clock::now();
    CU(h, cudaSetDevice(h->device));
    // workspace per clip, estimated on a batch large enough to include what only big batches allocate (the permuted copy
    // of the spectrogram for the good-first frame order)
    ainmf_params one = *p;
    one.batch = p->batch < 256 ? p->batch : 256;
    const size_t probe_ws = ainmf_workspace_bytes(h, &one);
    if (probe_ws == 0) return AINMF_ERR_INVALID;       // message set by make_plan
    const size_t per_clip_ws = (probe_ws + one.batch - 1) / one.batch;
    const long long N = p->n_samples;
    const size_t per_clip = per_clip_ws + 4 * sizeof(float) * (size_t)N + 256 +
                            (pcm ? 2 * sizeof(int16_t) * (size_t)N * (channels + 1) + 64 : 0);
    // Clips go through in chunks, two in flight: while chunk c is being fitted, chunk c+1 arrives on the copy-in stream and
    // the result of chunk c-1 leaves on the copy-out stream (x and y double-buffered, one workspace).  A chunk is at most 512
    // clips (and what device memory allows).
    // Chunk schedule.  A chunk costs a fixed ~8 ms (200 iterations x launch gaps, fill and drain of six kernels) whatever its
    // size, so few large chunks beat many small ones; what stays exposed is the copy-in of the first chunk and the copy-out
    // of the last.  So: a first chunk of n_sm clips (one clip per SM in every round of the persistent kernels, the smallest
    // size that still runs at ~95 % of the large-batch rate; 4.7 ms of copy-in at 10 s clips), then chunks as large as
    // memory and the 512-clip cap allow.  512 clips: [148, 364]; 4096: [148, 8 x 444, 396] (measured against [148, 148, 148, 68],
    // [148, 216, 148], [148, 296, 68], [222, 290], [296, 216], [74, 438]: profiles/r02e_e2e_chunk_schedules.txt).
    std::vector<long long> c_begin;          // first clip of every chunk, then the batch size
    long long chunk = 0;                     // clips in the largest chunk
    ainmf_params cp = *p;
    size_t ws = 0, o = 0, oX[2], oY[2], oNb[2], oEr[2], oNi[2], oWs = 0;
    size_t oPi[2] = {0, 0}, oPo[2] = {0, 0}, oPk[2] = {0, 0}, oPb = 0;     // 16-bit form: interleaved input, output, peaks; peak scratch
    auto schedule = [&](long long cap) {     // chunks of at most cap clips, and the scratch layout that holds them
        plan_host_chunks(p->batch, cap, h->n_sm, &c_begin);
        chunk = 0;
        for (size_t i = 0; i + 1 < c_begin.size(); ++i) if (c_begin[i + 1] - c_begin[i] > chunk) chunk = c_begin[i + 1] - c_begin[i];
        cp.batch = (int32_t)chunk;
        ws = ainmf_workspace_bytes(h, &cp);
        o = 0;
        auto take = [&](size_t bytes) { size_t r = o; o += al256(bytes); return r; };
        for (int i = 0; i < 2; ++i) {
            oX[i] = take(sizeof(float) * chunk * N); oY[i] = take(sizeof(float) * chunk * N);
            oNb[i] = take(sizeof(int) * chunk); oEr[i] = take(sizeof(float) * chunk); oNi[i] = take(sizeof(int) * chunk);
        }
        oWs = take(ws);
        if (pcm) {
            for (int i = 0; i < 2; ++i) {
                oPi[i] = take(sizeof(int16_t) * chunk * N * channels); oPo[i] = take(sizeof(int16_t) * chunk * N);
                oPk[i] = take(sizeof(float) * chunk);
            }
            oPb = take(sizeof(int) * chunk);
        }
    };
    if (max_device_bytes == 0) {
#ifndef AINMF_EMU
        // No cap given: plan for the 512-clip cap; only if the scratch block of an earlier call does not already hold that
        // layout is the free memory queried (cudaMemGetInfo is not cheap: 0.2 ms most of the time, 15-75 ms every fourth
        // call or so on the measured boxes -- up to a quarter of a 512-clip call).
        schedule(512);
        if (o > h->scratch_bytes) {
            size_t fr = 0, tot = 0;
            CU(h, cudaMemGetInfo(&fr, &tot));
            max_device_bytes = (size_t)((double)(fr + h->scratch_bytes) * 0.8);
        }
#else
        max_device_bytes = (size_t)1 << 30;
#endif
    }
    if (max_device_bytes != 0) {
        const long long cap = (long long)(max_device_bytes / per_clip);
        if (cap < 1) return fail(h, AINMF_ERR_WORKSPACE, "one clip needs %zu bytes of device memory", per_clip);
        schedule(cap);
    }
    void* scr;
    int rc = get_scratch(h, o, &scr);
    if (rc) return rc;
    char* base = (char*)scr;
    if (!h->st_in) {
        CU(h, cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
        CU(h, cudaStreamCreateWithFlags(&h->st_cmp, cudaStreamNonBlocking));
        CU(h, cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(h, cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            CU(h, cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
            CU(h, cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 4; ++i) CU(h, cudaEventCreateWithFlags(&h->ev_part[i], cudaEventDisableTiming));
    }
    void* pin;
    if ((rc = get_pinned(h, 4 * sizeof(int) * (size_t)p->batch, &pin))) return rc;
    int* stage_nb = (int*)pin;
    float* stage_er = (float*)(stage_nb + p->batch);
    int* stage_ni = (int*)(stage_er + p->batch);
    float* stage_pk = (float*)(stage_ni + p->batch);
    const long long n_chunks = (long long)c_begin.size() - 1;
    auto chunk_size = [&](long long c) { return (int)(c_begin[c + 1] - c_begin[c]); };
    const bool trace = getenv("AINMF_HOST_TRACE") != nullptr;   // host wall-clock of the phases, device time of every chunk's copies and fit
    std::vector<cudaEvent_t> tev, cev;
    auto mark = [&](std::vector<cudaEvent_t>& v, cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); v.push_back(e); } };
    auto copy_in = [&](long long c) -> int {
        const int j = (int)(c & 1);
        if (c >= 2) CU(h, cudaStreamWaitEvent(h->st_in, h->ev_done[j], 0));       // the fit of chunk c-2 has read this x buffer
        mark(cev, h->st_in);
        if (pcm) CU(h, cudaMemcpyAsync(base + oPi[j], pcm_in_host + c_begin[c] * N * channels, sizeof(int16_t) * (size_t)chunk_size(c) * N * channels, cudaMemcpyHostToDevice, h->st_in));
        else CU(h, cudaMemcpyAsync(base + oX[j], x_host + c_begin[c] * N, sizeof(float) * (size_t)chunk_size(c) * N, cudaMemcpyHostToDevice, h->st_in));
        mark(cev, h->st_in);
        CU(h, cudaEventRecord(h->ev_in[j], h->st_in));
        return 0;
    };
    const auto t_entry = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(); };
    std::vector<cudaEvent_t> oev;
    rc = AINMF_OK;
    if ((rc = copy_in(0))) return rc;
    for (long long c = 0; c < n_chunks && rc == AINMF_OK; ++c) {
        if (trace) fprintf(stderr, "[ainmf host] chunk %lld: enqueue starts at %.2f ms\n", c, since());
        const int j = (int)(c & 1), nb = chunk_size(c);
        const long long b0 = c_begin[c];
        if (c + 1 < n_chunks && (rc = copy_in(c + 1))) break;                      // enqueued before the fit blocks this thread at its polls
        cp.batch = nb;
        CU(h, cudaStreamWaitEvent(h->st_cmp, h->ev_in[j], 0));
        if (c >= 2) CU(h, cudaStreamWaitEvent(h->st_cmp, h->ev_out[j], 0));        // the result of chunk c-2 has left this y buffer
        mark(tev, h->st_cmp);
        if (pcm)      // load_damaged_data on the device: channel mean, peak, exact division (the x buffer is free: chunk c-2's fit is over)
            CU(h, launch_load_pcm16((const int16_t*)(base + oPi[j]), nb, N, channels, (float*)(base + oX[j]), (int*)(base + oPb), (float*)(base + oPk[j]), h->st_cmp));
        // the last chunk's copy-out is the exposed one: its waveform leaves in four parts behind the parts of the inverse STFT
        const bool sink = !pcm && c + 1 == n_chunks && nb >= 8 && !getenv("AINMF_HOST_NO_SINK");
        if (sink) {
            mark(oev, h->st_out);
            h->sink_y_host = y_host + b0 * N; h->sink_st = h->st_out;
        }
        rc = ainmf_inpaint(h, &cp, (const float*)(base + oX[j]), nullptr, nullptr, (float*)(base + oY[j]), nullptr,
                           (int*)(base + oNb[j]), nullptr, nullptr, (float*)(base + oEr[j]), (int*)(base + oNi[j]), base + oWs, ws, h->st_cmp);
        h->sink_y_host = nullptr; h->sink_st = nullptr;
        if (rc) break;
        if (pcm) CU(h, launch_store_pcm16((const float*)(base + oY[j]), (long long)nb * N, (int16_t*)(base + oPo[j]), h->st_cmp));   // save_result
        mark(tev, h->st_cmp);
        CU(h, cudaEventRecord(h->ev_done[j], h->st_cmp));
        CU(h, cudaStreamWaitEvent(h->st_out, h->ev_done[j], 0));
        if (!sink) mark(oev, h->st_out);
        if (pcm) {
            CU(h, cudaMemcpyAsync(pcm_out_host + b0 * N, base + oPo[j], sizeof(int16_t) * (size_t)nb * N, cudaMemcpyDeviceToHost, h->st_out));
            CU(h, cudaMemcpyAsync(stage_pk + b0, base + oPk[j], sizeof(float) * nb, cudaMemcpyDeviceToHost, h->st_out));
        } else if (!sink) {
            CU(h, cudaMemcpyAsync(y_host + b0 * N, base + oY[j], sizeof(float) * (size_t)nb * N, cudaMemcpyDeviceToHost, h->st_out));
        }
        // the per-clip scalars go through pinned staging: a copy into the caller's (pageable) arrays would block this thread
        // until the fit and the copy-out of this chunk are over, and nothing of the next chunk would be queued meanwhile
        CU(h, cudaMemcpyAsync(stage_nb + b0, base + oNb[j], sizeof(int) * nb, cudaMemcpyDeviceToHost, h->st_out));
        CU(h, cudaMemcpyAsync(stage_er + b0, base + oEr[j], sizeof(float) * nb, cudaMemcpyDeviceToHost, h->st_out));
        CU(h, cudaMemcpyAsync(stage_ni + b0, base + oNi[j], sizeof(int) * nb, cudaMemcpyDeviceToHost, h->st_out));
        mark(oev, h->st_out);
        CU(h, cudaEventRecord(h->ev_out[j], h->st_out));
    }
    // drain all three streams whatever happened, so that no copy is in flight when the caller's buffers go away
    const double t_enq = since();
    if (trace) fprintf(stderr, "[ainmf host] all chunks enqueued at %.2f ms\n", t_enq);
    cudaStreamSynchronize(h->st_in);
    cudaStreamSynchronize(h->st_cmp);
    if (trace) fprintf(stderr, "[ainmf host] compute stream drained at %.2f ms\n", since());
    const cudaError_t e_out = cudaStreamSynchronize(h->st_out);
    if (trace) {
        fprintf(stderr, "[ainmf host] copy-out drained at %.2f ms\n", since());
        auto at = [&](cudaEvent_t e) { float ms = 0.f; cudaEventElapsedTime(&ms, cev[0], e); return ms; };   // since the first copy-in began
        for (size_t c = 0; 2 * c + 1 < tev.size(); ++c)
            fprintf(stderr, "[ainmf host] chunk %zu on the device: copy-in %.2f-%.2f, fit %.2f-%.2f, copy-out %.2f-%.2f ms\n", c,
                    at(cev[2 * c]), at(cev[2 * c + 1]), at(tev[2 * c]), at(tev[2 * c + 1]),
                    2 * c + 1 < oev.size() ? at(oev[2 * c]) : 0.f, 2 * c + 1 < oev.size() ? at(oev[2 * c + 1]) : 0.f);
        for (auto* v : {&tev, &cev, &oev}) for (cudaEvent_t e : *v) cudaEventDestroy(e);
    }
    if (getenv("AINMF_HOST_TIMES"))      // one line per call: where the host spent it
        fprintf(stderr, "[ainmf host] setup %.2f ms, enqueue %.2f ms, drain %.2f ms\n",
                std::chrono::duration<double, std::milli>(t_entry - t_call).count(), t_enq, since() - t_enq);
    if (rc == AINMF_OK && e_out != cudaSuccess) return fail(h, AINMF_ERR_CUDA, "copy-out: %s", cudaGetErrorString(e_out));
    if (rc == AINMF_OK) {
        if (n_bad_host) memcpy(n_bad_host, stage_nb, sizeof(int) * (size_t)p->batch);
        if (err_host) memcpy(err_host, stage_er, sizeof(float) * (size_t)p->batch);
        if (n_iter_host) memcpy(n_iter_host, stage_ni, sizeof(int) * (size_t)p->batch);
        if (pcm && peak_host) memcpy(peak_host, stage_pk, sizeof(float) * (size_t)p->batch);
    }
    return rc;
}
}  // namespace

extern "C" {

int ainmf_inpaint_host(ainmf_handle h, const ainmf_params* p, const float* x_host, float* y_host, int32_t* n_bad_host,
                       float* err_host, int32_t* n_iter_host, size_t max_device_bytes) {
    if (!h) return AINMF_ERR_INVALID;
    if (!p || !x_host || !y_host) return fail(h, AINMF_ERR_INVALID, "params, x_host and y_host must not be NULL");
    return inpaint_host_impl(h, p, x_host, y_host, nullptr, 0, nullptr, nullptr, n_bad_host, err_host, n_iter_host, max_device_bytes);
}

int ainmf_host_chunk_schedule(int64_t batch, int64_t max_clips, int32_t n_sm, int32_t* sizes, int32_t max_sizes, int32_t* n_sizes) {
    if (batch < 1 || max_clips < 1 || n_sm < 1 || !sizes || !n_sizes || max_sizes < 1) return AINMF_ERR_INVALID;
    std::vector<long long> cb;
    plan_host_chunks(batch, max_clips, n_sm, &cb);
    const int n = (int)cb.size() - 1;
    *n_sizes = n;
    if (n > max_sizes) return AINMF_ERR_WORKSPACE;
    for (int i = 0; i < n; ++i) sizes[i] = (int32_t)(cb[i + 1] - cb[i]);
    return AINMF_OK;
}

int ainmf_inpaint_host_pcm16(ainmf_handle h, const ainmf_params* p, const int16_t* pcm_host, int32_t channels, int16_t* out_host,
                             float* peak_host, int32_t* n_bad_host, float* err_host, int32_t* n_iter_host, size_t max_device_bytes) {
    if (!h) return AINMF_ERR_INVALID;
    if (!p || !pcm_host || !out_host) return fail(h, AINMF_ERR_INVALID, "params, pcm_host and out_host must not be NULL");
    if (channels < 1 || channels > 8) return fail(h, AINMF_ERR_INVALID, "channels must be in [1,8], got %d", channels);
    return inpaint_host_impl(h, p, nullptr, nullptr, pcm_host, channels, out_host, peak_host, n_bad_host, err_host, n_iter_host, max_device_bytes);
}

unsigned long long ainmf_launch_count(void) { return ainmf::g_launch_count; }

int ainmf_profile(ainmf_handle h, int32_t enable, double* ms_out, int64_t* counts_out) {
    if (!h) return AINMF_ERR_INVALID;
    if (ms_out && counts_out) {
        double ms[PROF_KINDS];
        long long cn[PROF_KINDS];
        prof_collect(ms, cn);
        for (int i = 0; i < PROF_KINDS; ++i) { ms_out[i] = ms[i]; counts_out[i] = cn[i]; }
    }
    prof_enable(enable != 0);
    return AINMF_OK;
}

int ainmf_load_pcm16(ainmf_handle h, const int16_t* pcm, int32_t batch, int64_t n_samples, int32_t channels, float* x,
                     float* peak, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!pcm || !x || batch <= 0 || n_samples < 1 || channels < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_load_pcm16");
    CU(h, cudaSetDevice(h->device));
    void* scr;
    int rc = get_scratch(h, sizeof(int) * batch, &scr);
    if (rc) return rc;
    CU(h, launch_load_pcm16(pcm, batch, n_samples, channels, x, (int*)scr, peak, (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_store_pcm16(ainmf_handle h, const float* y, int64_t count, int16_t* pcm, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!y || !pcm || count < 0) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_store_pcm16");
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_store_pcm16(y, count, pcm, (cudaStream_t)stream));
    return AINMF_OK;
}

// ---- sample-level detectors / baselines of the sibling scripts, Part-0 post-processing ----------
int ainmf_find_main_gap(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, int64_t* span,
                        void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || !span || batch <= 0 || n_samples < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_find_main_gap");
    CU(h, cudaSetDevice(h->device));
    void* scr;
    int rc = get_scratch(h, gaps_work_bytes(batch, n_samples), &scr);
    if (rc) return rc;
    CU(h, launch_gap_span(x, n_samples, batch, n_samples, threshold, 0, scr, (long long*)span, nullptr, (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_find_gaps(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, int32_t min_len,
                    int64_t* runs, int32_t max_runs, int32_t* n_runs, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || !runs || !n_runs || batch <= 0 || n_samples < 1 || max_runs < 1 || min_len < 0)
        return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_find_gaps");
    CU(h, cudaSetDevice(h->device));
    void* scr;
    int rc = get_scratch(h, gaps_work_bytes(batch, n_samples), &scr);
    if (rc) return rc;
    CU(h, launch_gap_runs(x, n_samples, batch, n_samples, threshold, min_len, scr, (long long*)runs, max_runs, n_runs,
                          (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_linear_interp(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, float* y,
                        int64_t* n_damaged, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || !y || batch <= 0 || n_samples < 1) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_linear_interp");
    CU(h, cudaSetDevice(h->device));
    void* scr;
    int rc = get_scratch(h, gaps_work_bytes(batch, n_samples), &scr);
    if (rc) return rc;
    CU(h, launch_interp_fill(x, n_samples, batch, n_samples, threshold, scr, y, n_samples, (long long*)n_damaged, (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_apply_gaps(ainmf_handle h, float* x, int32_t batch, int64_t n_samples, const int64_t* starts, const int64_t* lens,
                     int32_t gaps_per_clip, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!x || !starts || !lens || batch <= 0 || n_samples < 1 || gaps_per_clip < 1)
        return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_apply_gaps");
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_apply_gaps(x, n_samples, batch, n_samples, (const long long*)starts, (const long long*)lens, gaps_per_clip,
                            (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_blend_boundaries(ainmf_handle h, const float* raw, const float* restored, int64_t n_samples, int64_t gap_start,
                           int64_t gap_end, int32_t blend_len, float* out, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!raw || !restored || !out || n_samples < 1 || blend_len < 2 || gap_start < blend_len || gap_end < gap_start ||
        gap_end + blend_len > n_samples)
        return fail(h, AINMF_ERR_INVALID, "ainmf_blend_boundaries needs blend_len >= 2 and blend_len <= gap_start <= gap_end <= n_samples - blend_len");
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_blend(raw, restored, n_samples, gap_start, gap_end, blend_len, out, (cudaStream_t)stream));
    return AINMF_OK;
}

int ainmf_snr_db(ainmf_handle h, const float* ref, const float* est, int64_t begin, int64_t end, double* snr_db, void* stream) {
    if (!h) return AINMF_ERR_INVALID;
    if (!ref || !est || !snr_db || begin < 0 || end < begin) return fail(h, AINMF_ERR_INVALID, "bad argument to ainmf_snr_db");
    CU(h, cudaSetDevice(h->device));
    void* scr;
    int rc = get_scratch(h, sizeof(double) * 512, &scr);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CU(h, launch_snr_sums(ref, est, begin, end, (double*)scr, s));
    double part[512];
    CU(h, cudaMemcpyAsync(part, scr, sizeof part, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    double num = 0.0, den = 0.0;
    for (int i = 0; i < 256; ++i) { num += part[2 * i]; den += part[2 * i + 1]; }
    *snr_db = 10.0 * log10(num / (den + 1e-10));
    return AINMF_OK;
}

}  // extern "C"

#include "sharded.inc"
