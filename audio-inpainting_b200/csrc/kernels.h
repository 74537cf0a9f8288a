// kernels.h -- host-callable launchers of the ainmf kernels (internal; the public surface is include/ainmf.h).
#pragma once
#include "common.cuh"

namespace ainmf {

// ---- tables owned by the handle (device memory) -------------------------------------------------
struct FftTables {
    int n_fft = 0;
    const float2* tw_half = nullptr;   // [n_fft/2]   exp(-2 pi i k / (n_fft/2))
    const float2* tw_full = nullptr;   // [n_fft/2+1] exp(-2 pi i k / n_fft)
    const float* window = nullptr;     // [n_fft]     analysis/synthesis window (float32)
    float win_sum = 0.f;               // float32 sum of the float32 window
};

struct StftGeom {
    long long N;     // samples per clip
    int n_fft, hop;
    int T, F, ldf;   // frames, bins, leading dimension of the [T][ldf] arrays
};

// ---- stft.cu ------------------------------------------------------------------------------------
// V[b][t][f] = |Z|, Z[b][t][f] = STFT (scipy.signal.stft semantics); x is [B][x_stride].
// t_offset: global index of local frame 0 (time-frame sharding: x points at the shard's first sample
// minus halo bookkeeping is done by the caller through x_origin): sample index of x[0] in the clip.
cudaError_t launch_stft(const float* x, long long x_stride, long long x_origin, long long x_avail, int B,
                        const StftGeom& g,
                        int t_begin, int t_count, const FftTables& tb, float* V, float2* Z,
                        long long vz_stride, cudaStream_t s);

// dynamic shared memory per block that the forward and / or inverse kernel needs for this geometry
size_t stft_smem_need(int n_fft, int hop, bool forward, bool inverse);

// y[b][n] for n in [n_begin, n_begin + n_count): inverse STFT with overlap-add of Z' where
// Z'[t] = bad[t] ? V[t] * Z[t]/|Z[t]| : Z[t].  Frames are local indices [0, t_count) that correspond
// to global frames [t_begin, t_begin + t_count); frames outside contribute nothing (caller adds halos).
// If n_bad[b] == 0 the kernel copies x (the reference returns the input untouched, main4_NMF_gap.py:54).
cudaError_t launch_istft(const float* V, const float2* Z, long long vz_stride, const unsigned char* bad,
                         long long bad_stride, const int* n_bad, const float* x, long long x_stride,
                         long long x_origin, int B, const StftGeom& g, int t_begin, int t_count,
                         const FftTables& tb, float* y, long long y_stride, long long n_begin,
                         long long n_count, int T_total, cudaStream_t s);

// ---- mask.cu ------------------------------------------------------------------------------------
// bad[b][c] for c in [0,T): column predicate of get_gap_mask / get_mask_from_signal.
cudaError_t launch_gap_mask(const float* x, long long x_stride, long long x_origin, long long x_avail,
                            int B, long long N, int hop, int t_begin, int T, float thr, int num, int den,
                            unsigned char* bad, long long bad_stride, cudaStream_t s);
// bad flags from an explicit column range (main4_NMF.py:74-76).
cudaError_t launch_range_mask(int B, int T, int col_start, int col_end, unsigned char* bad,
                              long long bad_stride, cudaStream_t s);
// ascending indices of set flags + count, one block per clip.
cudaError_t launch_compact(const unsigned char* bad, long long bad_stride, int B, int T, int* bad_idx,
                           long long idx_stride, int* n_bad, cudaStream_t s);

// ---- pcm.cu -------------------------------------------------------------------------------------
cudaError_t launch_load_pcm16(const int16_t* pcm, int B, long long N, int channels, float* x, int* peak_bits,
                              float* peak, cudaStream_t s);
cudaError_t launch_store_pcm16(const float* y, long long count, int16_t* pcm, cudaStream_t s);
// dst[b][c][r] = src[b][r][c] (r < rows, c < cols); with zero_pad the columns rows..ld_dst of dst are zeroed
cudaError_t launch_transpose_f32(const float* src, long long src_stride, int ld_src, int rows, int cols, float* dst,
                                 long long dst_stride, int ld_dst, int zero_pad, int B, cudaStream_t s);
cudaError_t launch_transpose_c64(const float2* src, long long src_stride, int ld_src, int rows, int cols, float2* dst,
                                 long long dst_stride, int ld_dst, int zero_pad, int B, cudaStream_t s);

// ---- gaps.cu: sample-level detectors / baselines of the sibling scripts and the Part-0 post-processing ----------
size_t gaps_work_bytes(int B, long long N);
cudaError_t launch_gap_span(const float* x, long long x_stride, int B, long long N, float thr, int inclusive, void* work,
                            long long* span, long long* n_gap, cudaStream_t s);
cudaError_t launch_gap_runs(const float* x, long long x_stride, int B, long long N, float thr, int min_len, void* work,
                            long long* runs, int max_runs, int* n_runs, cudaStream_t s);
cudaError_t launch_interp_fill(const float* x, long long x_stride, int B, long long N, float thr, void* work, float* y,
                               long long y_stride, long long* n_damaged, cudaStream_t s);
cudaError_t launch_apply_gaps(float* x, long long x_stride, int B, long long N, const long long* starts, const long long* lens,
                              int gaps_per_clip, cudaStream_t s);
cudaError_t launch_blend(const float* raw, const float* restored, long long N, long long gs, long long ge, int blend_len,
                         float* out, cudaStream_t s);
cudaError_t launch_snr_sums(const float* ref, const float* est, long long begin, long long end, double* sums, cudaStream_t s);

// ---- rng.cu: RandomState(seed).standard_normal on the device ------------------------------------------------------------
// The K*T_total normals of H0 (row-major (K, T_total); only frames [t_begin, t_begin + t_count) are kept, as Hn[K][t_count])
// followed by the F*K normals of W0 (Wn[F][K]) -- the order sklearn draws them in (_nmf.py:296-307).
cudaError_t launch_numpy_normals(uint32_t seed, int K, long long T_total, long long t_begin, long long t_count, float* Hn, int F, float* Wn,
                                 cudaStream_t s);

// ---- per-clip control block (device) ---------------------------------------------------------------
struct ClipState {
    int done;            // 1 once the stop rule fired (or nothing to do); kernels of the iteration skip the clip
    int n_iter;          // iterations performed (sklearn's n_iter_)
    int n_bad;           // number of bad frames
    int status;          // 0 ok, 1 no bad frames (input returned), 2 every frame bad (undefined in the reference)
    double viol_init;    // violation of iteration 1
    double viol_last;
    double sum_x;        // sum of the imputed spectrogram (for the init scale)
    float mean_x;        // mean of the imputed spectrogram
    float err;           // ||X - W H||_F
    double err_init;     // MU solver: error of the initial factors and of the previous check (sklearn tests
    double err_prev;     //            (previous_error - error) / error_at_init < tol every 10 iterations)
};

// ---- impute.cu ----------------------------------------------------------------------------------
struct ImputeWork {
    int frames_per_chunk = 0, n_chunks = 0;
    double* colsum = nullptr;    // [B][n_chunks][F]
    double* sums = nullptr;      // [B][F+1]: per-bin sums and the grand total (the buffer a time-sharded run all-reduces)
};
void impute_plan(int T, ImputeWork* wk);
size_t impute_work_bytes(int B, int F, const ImputeWork& wk);
void impute_carve(void* base, int B, int F, ImputeWork* wk);
// sums[b][f] = sum over frames not flagged in `excl` (null: all frames) of V[b][t][f]; sums[b][F] = total
cudaError_t launch_colsums(const float* V, long long v_stride, int ldf, int F, int T, int B,
                           const unsigned char* excl, long long excl_stride, const ImputeWork& wk, cudaStream_t s);
// fill[b][f] = sums[b][f] / (T_total - n_excl[b]); frames flagged in `bad` <- fill (in place: V becomes the
// NMF input X); state[b] initialised (n_bad, status, done).
cudaError_t launch_fill(float* V, long long v_stride, int ldf, int F, int T, int T_total, int B,
                        const unsigned char* bad, long long bad_stride, const int* n_bad, const int* n_excl,
                        float* fill /*[B][ldf]*/, ClipState* state, const ImputeWork& wk, cudaStream_t s);
// state[b].mean_x = sums[b][F] / (F*T_total); re-arms the iteration counters
cudaError_t launch_mean(int F, int T_total, int B, ClipState* state, const ImputeWork& wk, cudaStream_t s);
// W[b][f][k] = |avg_b * Wn[f][k]|, Ht[b][t][k] = |avg_b * Hn[k][t]|, avg_b = sqrt(mean_x/K) (float32);
// pad components k >= K are zero.  ($SP/sklearn/decomposition/_nmf.py:296-307)
cudaError_t launch_init_factors(const float* Wn /*[F][K]*/, const float* Hn /*[K][T_total]*/, int T_total,
                                int t_begin, int B, int F, int T, int K, int KP, const ClipState* state, float* W,
                                long long w_stride, float* Ht, long long h_stride, cudaStream_t s);
// user factors W0[b][F][K], H0[b][K][T] -> padded internal layout
cudaError_t launch_pack_factors(const float* W0, const float* H0, int B, int F, int T, int K, int KP, float* W,
                                long long w_stride, float* Ht, long long h_stride, cudaStream_t s);
// internal -> user layout W[b][F][K], H[b][K][T]
cudaError_t launch_unpack_factors(const float* W, long long w_stride, const float* Ht, long long h_stride, int B,
                                  int F, int T, int K, int KP, float* Wout, float* Hout, cudaStream_t s);
// copy selected ClipState fields into plain arrays (any may be null)
cudaError_t launch_export_state(const ClipState* st, int B, int* n_bad, int* n_iter, float* err, int* status,
                                cudaStream_t s);
// frame permutation for the fit: good frames first (perm[b][t'] = original frame), row gather / scatter
cudaError_t launch_invert_flags(const unsigned char* bad, long long stride, int B, int T, unsigned char* good, cudaStream_t s);
cudaError_t launch_build_perm(const int* good_idx, const int* n_good, const int* bad_idx, int B, int T, int* perm, cudaStream_t s);
cudaError_t launch_gather_rows(const float* src, long long src_stride, float* dst, long long dst_stride, int ld, const int* perm,
                               int B, int T, const int* limit, cudaStream_t s);
cudaError_t launch_scatter_rows(const float* src, long long src_stride, float* dst, long long dst_stride, int ld, const int* perm,
                                int B, int T, cudaStream_t s);
// state for a fit on a caller-provided X (no imputation stage): done=0, n_iter=0
cudaError_t launch_reset_state(ClipState* st, int B, cudaStream_t s);

// ---- profiling (CUDA events around the kernels of the iteration; off by default) -----------------------
enum ProfKind { PROF_GRAM_H = 0, PROF_XHT, PROF_W_SWEEP, PROF_GRAM_W, PROF_H_STEP, PROF_STOP, PROF_KINDS };
void prof_enable(bool on);
void prof_begin(int kind, cudaStream_t s);
void prof_end(int kind, cudaStream_t s);
// synchronises the recorded events, adds the elapsed milliseconds / launch counts per kind, then resets
void prof_collect(double* ms /*[PROF_KINDS]*/, long long* counts /*[PROF_KINDS]*/);

// ---- nmf_cd.cu ----------------------------------------------------------------------------------
struct NmfProblem {
    int B, T, F, ldf, KP;
    int solver = 0;                     // 0 = coordinate descent (the reference's), 1 = multiplicative update (Frobenius), 2 = MU (Kullback-Leibler)
    float tol;
    float* Xt; long long x_stride;      // [B][T][ldf]
    float* W; long long w_stride;       // [B][F][KP]
    float* Ht; long long h_stride;      // [B][T][KP]
    // good-first frame order (tensor-core path, one fit): frames t >= t_good[b] are bad = all equal to fill[b][:]; null: off
    const int* t_good = nullptr;
    const float* fill = nullptr; long long fill_stride = 0;
    ClipState* state;                   // [B]
};
// TMA tensor maps of the tensor-core path (each an opaque 128-byte CUtensorMap; built on the host per problem)
struct alignas(64) TcMapBlob { unsigned char b[128]; };
struct TcMaps { TcMapBlob mapX, mapWt, mapWtLo, mapXs, mapHs, mapHmn, mapHk, mapG, mapGlo; };

constexpr int kSweepScalars = 136;   // per block of 8 coordinates: G diagonal block 8x8, look-ahead block 8x8, 1/diag (8)

constexpr int kWSideTail = 32;          // rows beyond ROWS that the last W-side block of a clip may take (one lane per row shape)
struct NmfWork {
    int h_bm = 64, nW = 0, nH = 0, xht_splits = 1, gram_max_blocks = 64, w_lanes = 1, w_rows = 128, finish_gpb = 1;
    // tensor-core path (tcgen05: tf32 main term + bf16 cross terms) for the two V-sized contractions; KP in {64,128}
    int use_tc = 0, tc_splits = 1, tc_fps = 0, tc_mtiles = 0;
    float *tc_Wt = nullptr, *tc_WtLo = nullptr;   // [B][KP][ldf]: W transposed (tf32 main-term operand), and its bf16 cross-term operand (tc::cross_pack8)
    float* tc_GLo = nullptr;                      // [B][KP][KP]: bf16 cross-term operand of W^T W
    float* tc_blobs = nullptr;                    // [B][KP/8][16*KP]: per-block update operands of the sweep (nmf_ts.cu: g_prep_kernel)
    float* tc_scal = nullptr;                     // [B][KP/8][kSweepScalars]: per-block sweep scalars
    float* tc_vpartial = nullptr;                 // [B][nW][KP]: per-block shares of fill^T.W (good-first frame order)
    float* tc_hbad = nullptr;                     // [B][KP]: sum of the bad frames' rows of Ht (good-first frame order)
    float* tc_hbad_part = nullptr; int hbad_blocks = 1;   // [B][hbad_blocks][KP]: shares of tc_hbad
    float* tc_vfill = nullptr;                    // [B][KP]: fill^T.W, the X^T.W row of every bad frame (good-first frame order)
    const TcMaps* tc = nullptr;
    float *HHt = nullptr, *WtW = nullptr, *gram_partial = nullptr, *xht_partial = nullptr;
    float *violW = nullptr, *violH = nullptr;
    // small problems on the FFMA path (the early-stopping refit loop of main4_NMF.py): |pg| of every (row, coordinate) so that
    // the stop rule can add them in sklearn's own order and precision when the decision is close (stop_kernel)
    int exact_viol = 0;
    float *pgW = nullptr, *pgH = nullptr;   // [B][F][KP], [B][T][KP]
    unsigned* counters = nullptr;       // must be zero before the first iteration
    double* err_partial = nullptr;
    // time-frame-sharded mode only (B == 1); null otherwise
    float* xht_reduced = nullptr;       // [F][KP]: local X.Ht summed over the splits = first part of the all-reduce buffer
    double* h_viol_sum = nullptr;       // [B]: local H-side violation, all-reduced by the caller before the stop rule
    float* h_viol_pack = nullptr;       // [B][4]: the same as hi/lo floats, for a caller that sends it with the next float all-reduce
    // multiplicative-update solver
    float* xtw = nullptr;               // [B][T][KP]: X^T.W (the MU numerator of the H update)
    unsigned char* zero_flags = nullptr;  // [B][zero_stride] zeros: error evaluation that must not overwrite frames
    long long zero_stride = 0;
    int want_mu = 0;                    // set before nmf_work_bytes / nmf_carve: 1 = MU (Frobenius), 2 = MU (Kullback-Leibler)
    char* coop_scratch = nullptr;       // 4 KB: grid-barrier word + per-CTA violation shares of the one-clip cooperative fit
    float* kl_sums = nullptr;           // MU-KL: [2][B][KP] row sums of H, column sums of W (the denominators)
    double* kl_err = nullptr;           // MU-KL: [B] final divergence (kept across nmf_finalize, which reports the Frobenius error)
    // side stream for the small reductions that depend only on Ht (hbad): they run next to the X.Ht kernel instead of
    // after it.  Null: everything on the caller's stream.
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
enum { NMF_PHASE_PARTIALS = 1, NMF_PHASE_UPDATE = 2, NMF_PHASE_STOP = 4,
       NMF_PHASE_STOP_PACKED = 8 };   // with STOP: the H-side violation is read from wk.h_viol_pack (summed hi + lo floats)
void nmf_plan(int B, int T, int F, int KP, int n_sm, NmfWork* wk);
size_t nmf_work_bytes(int B, int T, int F, int KP, const NmfWork& wk);
// tensor-core path: build the TMA maps for problem `p` (after nmf_carve) and attach them; 0 on success
int nmf_tc_setup(const NmfProblem& p, NmfWork* wk, TcMaps* maps);
cudaError_t nmf_tc_half1(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);   // X.Ht partials + HHt
cudaError_t nmf_tc_hstep(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);   // Wt split, X^T.W + H sweep
cudaError_t nmf_ts_hstep(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);   // persistent TMEM-operand H step (nmf_ts.cu)
cudaError_t nmf_ts_half1(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);   // persistent TMEM-operand X.Ht + Gram (nmf_ts.cu)
void nmf_carve(void* base, int B, int T, int F, int KP, NmfWork* wk);
cudaError_t nmf_cd_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s);
// multiplicative update, Frobenius loss (sklearn solver='mu': $SP/sklearn/decomposition/_nmf.py:536-549,615-624,
// 633-635,701-721,867-879).  nmf_mu_begin evaluates the error of the initial factors; nmf_mu_iterate does one
// W and H update and, every 10th iteration when tol > 0, the convergence test.
cudaError_t nmf_mu_begin(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);
cudaError_t nmf_mu_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s);
cudaError_t nmf_mu_tick(const NmfProblem& p, int it, cudaStream_t s);     // n_iter = it for the clips still iterating
cudaError_t nmf_mu_stop(const NmfProblem& p, int it, cudaStream_t s);     // the every-10th-iteration test on state.err
// nmf_mukl.cu: multiplicative update, generalised Kullback-Leibler divergence (sklearn solver='mu',
// beta_loss='kullback-leibler': _nmf.py:551-626, 636-721, 129-154, 862-879), one fused kernel per half-step.
// nmf_mukl_error writes sqrt(2 D_KL) to state.err (keep = false) or to wk.kl_err (keep = true: the final value, computed
// before nmf_finalize replaces frames; nmf_mukl_set_err puts it back into state.err afterwards).
cudaError_t nmf_mukl_begin(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);
cudaError_t nmf_mukl_iterate(const NmfProblem& p, const NmfWork& wk, int it, cudaStream_t s);
cudaError_t nmf_mukl_error(const NmfProblem& p, const NmfWork& wk, bool keep, cudaStream_t s);
cudaError_t nmf_mukl_set_err(const NmfProblem& p, const NmfWork& wk, cudaStream_t s);
// the same iteration in pieces, so that a collective can be placed between them (time-sharded mode):
// PARTIALS: HHt and X.Ht of the local frames; UPDATE: W sweep, WtW, fused X^T.W + H sweep; STOP: the stop rule
cudaError_t nmf_cd_phase(const NmfProblem& p, const NmfWork& wk, int it, int phases, cudaStream_t s);
// nmf_wside.cu: the W half-step (w_side_kernel + w_finish_kernel); S = number of X.Ht partial buffers to sum
cudaError_t launch_w_side(const NmfProblem& p, const NmfWork& wk, int S, cudaStream_t s);
cudaError_t launch_set_err(ClipState* st, int B, const double* err_sq, cudaStream_t s);
// ---- nmf_coop.cu: one spectrogram's whole fit as one cooperative persistent launch (all SMs, grid barriers between the
// four phases of an iteration, stop rule on the device).  On return W / Ht hold the factors, state n_iter / viol / done.
bool nmf_coop_eligible(const NmfProblem& p, const NmfWork& wk, int n_sm);
cudaError_t nmf_coop_fit(const NmfProblem& p, const NmfWork& wk, int max_iter, int n_sm, cudaStream_t s);
// ---- nmf_small.cu: a small spectrogram's whole fit -- and main4_NMF.py's chain of n_outer refits -- in one launch, one CTA
// per clip, everything in shared memory (seeded initial factors).  bad: frames replaced by (W H)
// after each fit.  On return Xt holds the restored frames, W / Ht the last fit's factors, state n_iter / err.
bool nmf_small_eligible(int F, int T, int K);      // K = the rank itself, not the padded one
cudaError_t nmf_small_refit(const NmfProblem& p, int K, const float* Wn, const float* Hn, const unsigned char* bad, long long bad_stride,
                            int n_outer, int max_iter, cudaStream_t s);
// err = ||X - W H||_F into state[b].err, then bad frames of Xt <- (W H) frames
cudaError_t nmf_finalize(const NmfProblem& p, const NmfWork& wk, const unsigned char* bad, long long bad_stride,
                         cudaStream_t s, double* err_sq = nullptr);

}  // namespace ainmf
