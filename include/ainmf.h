/*
 * ainmf.h -- C ABI of libainmf.so: B200-native NMF spectrogram inpainting.
 *
 * The reference (conniemessi/Audio-Inpainting) has no FFI; its boundary for this path is the body of
 *   NMFFairGapInpainter.restore / get_gap_mask        main4_NMF_gap.py:42-72 / :28-40
 *   NMFFairInpainter.restore / get_mask_from_signal   main4_NMF_mask.py:47-77 / :28-45
 *   SpectralInpainter.restore_with_nmf                main4_NMF.py:62-112
 * i.e. normalised float32 waveform in -> restored float32 waveform out, through three third-party calls
 * (scipy.signal.stft, sklearn.decomposition.NMF.fit_transform, scipy.signal.istft).  Each entry point below
 * names the reference lines it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the name ends in _host; the caller owns all buffers;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*) or on a side stream of the handle that forks
 *     from and joins back into `stream` inside the call (event-ordered, invisible to the caller: whatever is enqueued on
 *     `stream` after the call runs after all of it); nothing synchronises unless stated;
 *   - return value: 0 on success, negative ainmf_status otherwise; text via ainmf_last_error();
 *   - no CPU fallback exists: without a CUDA device ainmf_create fails;
 *   - threads and streams: a handle owns scratch memory, pinned staging, cached FFT / window / N(0,1) tables and the
 *     pipeline streams of ainmf_inpaint_host, none of it locked.  Calls on ONE handle must not overlap in time (serialise
 *     them, or give each thread its own handle), and consecutive calls on one handle should use one stream -- a caller
 *     that changes streams between calls must order those streams itself, because scratch is reused from call to call;
 *   - layouts are the reference's: spectrogram-shaped outputs are (F, T) row-major exactly like the
 *     arrays `signal.stft` returns and `NMF` consumes ("W" is (F, K), "H" is (K, T)), EXCEPT the *_tf
 *     entry points, which use the library's internal frame-major [T][ldf] layout and avoid transposes.
 */
#ifndef AINMF_H
#define AINMF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ainmf_context* ainmf_handle;

enum ainmf_status {
    AINMF_OK = 0,
    AINMF_ERR_INVALID = -1,     /* bad argument (message says which) */
    AINMF_ERR_CUDA = -2,        /* CUDA runtime error */
    AINMF_ERR_NO_DEVICE = -3,   /* no usable sm_100 device */
    AINMF_ERR_WORKSPACE = -4,   /* workspace too small */
    AINMF_ERR_ALL_BAD = -5,     /* every frame of a clip is flagged: mean of empty set (reference: NaN) */
    AINMF_ERR_COMM = -6         /* NCCL error */
};

enum ainmf_solver {
    AINMF_SOLVER_CD = 0,        /* sklearn solver='cd', beta_loss='frobenius' -- what the reference runs */
    AINMF_SOLVER_MU = 1,        /* multiplicative update, Frobenius (sklearn solver='mu') */
    AINMF_SOLVER_MU_KL = 2      /* multiplicative update, generalised Kullback-Leibler divergence (sklearn solver='mu',
                                   beta_loss='kullback-leibler'): the masked-ratio form X / (W H); err = sqrt(2 D_KL) */
};

/* Parameters of one inpainting problem.  Defaults of the reference in brackets. */
typedef struct ainmf_params {
    int32_t batch;          /* B independent clips */
    int64_t n_samples;      /* N samples per clip */
    int32_t n_fft;          /* nperseg [1024 gap/mask, 512 part0]; power of two, 64..4096 */
    int32_t hop;            /* n_fft - noverlap [256 / 128]; must divide n_fft, n_fft/hop <= 8 */
    int32_t rank;           /* n_components K [40]; 1..128 */
    int32_t max_iter;       /* [200] */
    float tol;              /* [1e-4] */
    int32_t solver;         /* ainmf_solver [CD] */
    uint32_t seed;          /* random_state [42 gap/mask, 0 part0]; used when W0/H0 are null */
    float threshold;        /* |x| < threshold marks a silent sample [1e-4 gap, 0.01 mask] */
    int32_t frac_num;       /* column is bad iff silent fraction > frac_num/frac_den [9/10 gap, 4/5 mask] */
    int32_t frac_den;
    int32_t col_start;      /* >= 0: bad frames are [col_start, col_end) and the fill is the mean of the */
    int32_t col_end;        /*       frames before col_start (main4_NMF.py:74-81); < 0: detect from x  */
    int32_t n_outer;        /* number of refits, each followed by bad frames <- W.H [1; 50 in main4_NMF.py:86-90] */
} ainmf_params;

/* ---- handle ------------------------------------------------------------------------------------------ */
int ainmf_create(ainmf_handle* out, int device);
int ainmf_destroy(ainmf_handle h);
const char* ainmf_last_error(ainmf_handle h);   /* h may be NULL: error of the last failed ainmf_create */
const char* ainmf_version(void);
void ainmf_params_default(ainmf_params* p);     /* the constants of main4_NMF_gap.py */

/* ---- geometry ---------------------------------------------------------------------------------------- */
/* Frame count T, bins F = n_fft/2+1 and internal leading dimension ldf of scipy.signal.stft(x, nperseg=n_fft,
 * noverlap=n_fft-hop) with its defaults boundary='zeros', padded=True ($SP/scipy/signal/_spectral_py.py:2240-2247). */
int ainmf_stft_geometry(int64_t n_samples, int32_t n_fft, int32_t hop, int32_t* T, int32_t* F, int32_t* ldf);
int32_t ainmf_padded_rank(int32_t rank);        /* 32, 64 or 128 */

/* ---- initial factors ---------------------------------------------------------------------------------- */
/* out[i] = float32(numpy.random.RandomState(seed).standard_normal(n)[i]) on the device: MT19937 + the legacy polar Gaussian,
 * the stream sklearn's init='random' draws its factors from ($SP/sklearn/decomposition/_nmf.py:296-307: H0 = |avg * N((K, T))|
 * first, then W0 = |avg * N((F, K))|, avg = sqrt(mean(X) / K)).  ainmf_nmf_fit / ainmf_inpaint use it internally when W0/H0
 * are NULL; it is exported for callers that build W0/H0 themselves.  One thread block walks the stream: ~2 ns per normal. */
int ainmf_standard_normal(ainmf_handle h, uint32_t seed, int64_t n, float* out, void* stream);

/* ---- window -------------------------------------------------------------------------------------------- */
/* The `window` argument of scipy.signal.stft / istft (the reference passes none: main4_NMF_gap.py:47,71 run scipy's default,
 * the periodic Hann window, which is what every call uses until this is called).  window_host: n_fft float32 values on the
 * HOST (what np.asarray(window, float32) holds), or NULL to go back to the periodic Hann window.  It applies to every later
 * ainmf_stft / ainmf_istft / ainmf_inpaint* call of this handle with that n_fft; the same window analyses and synthesises,
 * as in scipy, so it must satisfy NOLA for the hop in use.  Call it while no work of this handle is in flight. */
int ainmf_set_window(ainmf_handle h, int32_t n_fft, const float* window_host);

/* ---- stage entry points (device buffers) --------------------------------------------------------------- */
/* signal.stft + np.abs (main4_NMF_gap.py:47-48).  mag_ft [B][F][T] float, Z_ft [B][F][T] complex64 (re,im). Either
 * may be NULL.  The window is periodic Hann, scipy's default, unless ainmf_set_window installed another. */
int ainmf_stft(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, int32_t n_fft, int32_t hop,
               float* mag_ft, float* Z_ft, void* stream);

/* get_gap_mask / get_mask_from_signal (main4_NMF_gap.py:28-40, main4_NMF_mask.py:28-45).
 * bad [B][n_frames] uint8 flags; bad_idx [B][n_frames] int32 ascending indices (first n_bad[b] valid); n_bad [B]. */
int ainmf_gap_mask(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, int32_t hop, int32_t n_frames,
                   float threshold, int32_t frac_num, int32_t frac_den, uint8_t* bad, int32_t* bad_idx,
                   int32_t* n_bad, void* stream);

/* NMF(n_components=K, init='custom'|'random', max_iter, tol).fit_transform(X) (main4_NMF_gap.py:62-64;
 * $SP/sklearn/decomposition/_nmf.py:399-518, _cdnmf_fast.pyx:8-38).  X_ft [B][F][T] non-negative.
 * W0 [B][F][K], H0 [B][K][T]: initial factors, or both NULL to draw them from `seed` exactly as
 * sklearn's init='random' does.  Outputs W [B][F][K], H [B][K][T], err [B] (reconstruction_err_),
 * n_iter [B] (n_iter_); any output may be NULL. */
int ainmf_nmf_fit(ainmf_handle h, const float* X_ft, int32_t batch, int32_t F, int32_t T, int32_t rank,
                  int32_t max_iter, float tol, int32_t solver, uint32_t seed, const float* W0, const float* H0,
                  float* W, float* H, float* err, int32_t* n_iter, void* stream);

/* signal.istft + [:N] (main4_NMF_gap.py:71-72).  Z_ft [B][F][T] complex64; y [B][N]. */
int ainmf_istft(ainmf_handle h, const float* Z_ft, int32_t batch, int32_t T, int32_t n_fft, int32_t hop,
                int64_t n_samples, float* y, void* stream);

/* ---- the whole path --------------------------------------------------------------------------------- */
/* Bytes of device workspace ainmf_inpaint needs for `p` (0 on invalid parameters). */
size_t ainmf_workspace_bytes(ainmf_handle h, const ainmf_params* p);

/* restore() of main4_NMF_gap.py:42-72 / main4_NMF_mask.py:47-77, or restore_with_nmf() of main4_NMF.py:62-95
 * when p->col_start >= 0 (without _blend_boundaries, which needs the ground truth).
 *   x [B][N] normalised corrupted waveform -> y [B][N] restored waveform (y = x for clips with no bad frame).
 * Optional outputs (NULL to skip): bad_idx [B][T] int32, n_bad [B], W [B][F][K], H [B][K][T], err [B], n_iter [B]
 * (of the last refit).  W0/H0 as in ainmf_nmf_fit.  Synchronises `stream` only to poll the stop flags
 * (every few iterations) when p->tol > 0. */
int ainmf_inpaint(ainmf_handle h, const ainmf_params* p, const float* x, const float* W0, const float* H0, float* y,
                  int32_t* bad_idx, int32_t* n_bad, float* W, float* H, float* err, int32_t* n_iter,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Same with HOST buffers: stages x through pinned memory, runs ainmf_inpaint in clip chunks sized to
 * `max_device_bytes` (0: 80 % of free memory), copies y (and the optional outputs) back, and synchronises.
 * This is the call bench.py times for the end-to-end figure. */
int ainmf_inpaint_host(ainmf_handle h, const ainmf_params* p, const float* x_host, float* y_host, int32_t* n_bad_host,
                       float* err_host, int32_t* n_iter_host, size_t max_device_bytes);

/* The same from file samples to file samples, the way the scripts run (main4_NMF_gap.py:17-26 load_damaged_data ->
 * :42-72 restore -> :74-78 save_result): pcm_host [B][N][channels] interleaved int16 as scipy.io.wavfile.read returns
 * it -> channel mean -> x / max|x| (ainmf_load_pcm16, bit-exact) -> ainmf_inpaint -> clip * 32767 truncated
 * (ainmf_store_pcm16) -> out_host [B][N] int16 = what wavfile.write receives.  peak_host [B] (optional): max|x| before
 * the normalisation.  Half the bytes of the float32 form cross PCIe in each direction; the conversions run on the device
 * inside the chunk pipeline.  Pinned host buffers let the copies overlap the fits. */
int ainmf_inpaint_host_pcm16(ainmf_handle h, const ainmf_params* p, const int16_t* pcm_host, int32_t channels,
                             int16_t* out_host, float* peak_host, int32_t* n_bad_host, float* err_host,
                             int32_t* n_iter_host, size_t max_device_bytes);

/* The chunk schedule the two host entry points use for `batch` clips when at most `max_clips` fit the device at once
 * (host logic only, no device needed): sizes[0..*n_sizes) clips per chunk, in order.  A first chunk of n_sm clips (short
 * exposed copy-in), then multiples of n_sm up to min(max_clips, 512), what is left with the last chunk. */
int ainmf_host_chunk_schedule(int64_t batch, int64_t max_clips, int32_t n_sm, int32_t* sizes, int32_t max_sizes,
                              int32_t* n_sizes);

/* ---- front/back end of the scripts ------------------------------------------------------ */
/* load_damaged_data (main4_NMF_gap.py:21-24): int16 [B][N][channels] -> mono mean -> float32 -> x / max|x| (true
 * division).  peak [B] optional output. */
int ainmf_load_pcm16(ainmf_handle h, const int16_t* pcm, int32_t batch, int64_t n_samples, int32_t channels, float* x,
                     float* peak, void* stream);
/* save_result (main4_NMF_gap.py:76-77): clip to [-1,1], * 32767, truncate toward zero. */
int ainmf_store_pcm16(ainmf_handle h, const float* y, int64_t count, int16_t* pcm, void* stream);

/* ---- callers / baselines either side of the NMF path ----------------------------------- */
/* find_main_gap (main3_AR_text_gap.py:34-49): span [B][2] int64 = {first, last + 1} of the samples with
 * |x| < threshold, {-1, -1} when there is none. */
int ainmf_find_main_gap(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, int64_t* span,
                        void* stream);
/* find_gaps (main3_AR_text_mask.py:30-52): maximal runs [s, e) of |x| < threshold with e - s > min_len (100 in the
 * script), ascending; runs [B][max_runs][2] int64, n_runs [B] (the count found, which may exceed max_runs). */
int ainmf_find_gaps(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, int32_t min_len,
                    int64_t* runs, int32_t max_runs, int32_t* n_runs, void* stream);
/* linear_interp_part1.py:52-75: valid = |x| > threshold; y = x on valid samples and np.interp over the valid ones
 * elsewhere (float64 arithmetic, end values outside the valid range).  n_damaged [B] int64 optional. */
int ainmf_linear_interp(ainmf_handle h, const float* x, int32_t batch, int64_t n_samples, float threshold, float* y,
                        int64_t* n_damaged, void* stream);
/* Fixture producers (generate_part1_data.py:44-46 after create_random_mask, generate_part2_data.py:36-43): zero
 * x[b][s : s + l) for the clip's gaps_per_clip gaps; starts/lens [B][gaps_per_clip] int64 on the device (drawn by the
 * caller, e.g. with np.random.seed(clip)); entries with a negative start or non-positive length are skipped. */
int ainmf_apply_gaps(ainmf_handle h, float* x, int32_t batch, int64_t n_samples, const int64_t* starts, const int64_t* lens,
                     int32_t gaps_per_clip, void* stream);
/* _blend_boundaries (main4_NMF.py:114-126): ground truth outside [gap_start, gap_end), restored inside, linear
 * cross-fades of blend_len (50) samples either side (np.linspace ramp, float64 arithmetic). */
int ainmf_blend_boundaries(ainmf_handle h, const float* raw, const float* restored, int64_t n_samples, int64_t gap_start,
                           int64_t gap_end, int32_t blend_len, float* out, void* stream);
/* 10 log10(sum ref^2 / (sum (ref - est)^2 + 1e-10)) over samples [begin, end) (main4_NMF.py:99-110); sums in double;
 * synchronises the stream, result on the host. */
int ainmf_snr_db(ainmf_handle h, const float* ref, const float* est, int64_t begin, int64_t end, double* snr_db, void* stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------ */
/* Number of kernels this library has launched in this process (monotonic). */
unsigned long long ainmf_launch_count(void);
/* Kernel timing of the NMF iteration with CUDA events on the launching stream.  Reads (and clears) the time
 * accumulated since the last call into ms_out[6] / counts_out[6] -- order: gram(Ht) [FFMA path only], X.Ht partials
 * (+ Gram of Ht on the tensor-core path), W side (sweep, W^T W, operands of the H step), gram(W) [MU solver only],
 * fused X^T.W + H sweep, stop rule -- then switches recording on or off.  ms_out/counts_out may be NULL. */
int ainmf_profile(ainmf_handle h, int32_t enable, double* ms_out, int64_t* counts_out);

/* ---- time-frame sharding of one long signal ------------------------------------------------- */
/* Join `nranks` handles (one per GPU / process) into a group.  unique_id is the 128-byte NCCL id produced by
 * ainmf_comm_unique_id on rank 0 and distributed by the caller (torch.distributed broadcast). */
int ainmf_comm_unique_id(uint8_t id_out[128]);
int ainmf_comm_init(ainmf_handle h, const uint8_t unique_id[128], int32_t rank, int32_t nranks);
/* Which transport carries the per-iteration sums of ainmf_inpaint_sharded: 0 = no group / one rank, 1 = ncclAllReduce
 * or the caller's callbacks, 2 = peer mailboxes (every rank writes its F*K + K*K partial sums straight into a slot on
 * every other rank over NVLink -- CUDA IPC mappings made at the first sharded call -- and sums the slots in rank order
 * in a kernel of its own; no library collective between the kernels of an iteration).  Mailboxes are the default with
 * the NCCL group; AINMF_PEER_EXCHANGE=0 in the environment keeps ncclAllReduce. */
int ainmf_comm_transport(ainmf_handle h);
/* Alternative transport: the caller supplies the two collectives (used by the CPU test-suite over gloo).
 * dtype: 0 float32, 1 float64, 2 int32; op: 0 sum, 1 max; peers < 0 mean "none"; return 0 on success. */
typedef int (*ainmf_allreduce_fn)(void* user, void* buf, size_t count, int dtype, int op, void* stream);
typedef int (*ainmf_sendrecv_fn)(void* user, const void* sendbuf, int send_peer, void* recvbuf, int recv_peer,
                                 size_t n_floats, void* stream);
int ainmf_comm_set_callbacks(ainmf_handle h, int32_t rank, int32_t nranks, ainmf_allreduce_fn allreduce,
                             ainmf_sendrecv_fn sendrecv, void* user);
/* Frames [t_begin, t_end) of a T-frame signal owned by `rank`, and the sample range of x that rank needs
 * (its frames' support plus the mask/OLA halo), clipped to [0, N). */
int ainmf_shard_plan(int64_t n_samples, int32_t n_fft, int32_t hop, int32_t rank, int32_t nranks, int32_t* t_begin,
                     int32_t* t_end, int64_t* x_begin, int64_t* x_end, int64_t* y_begin, int64_t* y_end);
size_t ainmf_sharded_workspace_bytes(ainmf_handle h, const ainmf_params* p);
/* One rank's part of the whole path on a signal split by time frames: x_local = x[x_begin:x_end) of the
 * plan, y_local = y[y_begin:y_end).  W is replicated, H/V are sliced; per iteration the F*K + K*K partial sums of
 * the W half-step (and one violation scalar) are all-reduced over NVLink.  batch must be 1. */
int ainmf_inpaint_sharded(ainmf_handle h, const ainmf_params* p, const float* x_local, float* y_local, int32_t* n_bad,
                          float* W, float* H_local, float* err, int32_t* n_iter, void* workspace, size_t workspace_bytes,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AINMF_H */
