/* pnp_b200.h -- C ABI of libpnp_b200.so: the sm_100a kernels behind the PnP iteration hot path.
 *
 * The reference (vmonardo/pnp-svrg) is pure Python and has no FFI layer; its boundary is the
 * duck-typed protocol  problem.{grad_full,grad_stoch,select_mb,PSNR} / denoiser.denoise  that
 * algorithms/pnp_*.py call.  Each entry point below names the reference line(s) it replaces.
 * The Python host side (pnp_svrg_b200/) binds these with ctypes; INTEGRATION.md shows the stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - images are float32 and TRANSPOSED: [batch][W lines][H samples], line c = original column c;
 *   - `stream` is a cudaStream_t passed as void*; nothing here touches the default stream
 *     unless the caller passes it, nothing synchronises unless stated;
 *   - every function returns 0 on success or a negative pnp_status; pnp_last_error() gives text;
 *   - H and W must be powers of two in [32, 4096].
 */
#ifndef PNP_B200_H
#define PNP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    PNP_OK = 0,
    PNP_ERR_ARG = -1,       /* bad argument (size not supported, null pointer, ...) */
    PNP_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    PNP_ERR_NOT_INIT = -3,  /* pnp_init() has not been called on this device */
    PNP_ERR_UNSUPPORTED = -4 /* this entry point cannot handle the given size; use the documented alternative */
} pnp_status;

/* Library / device set-up: uploads the twiddle table and raises the dynamic shared-memory
 * limits of the FFT kernels on the CURRENT device.  Idempotent. */
int pnp_init(void);
const char* pnp_last_error(void);
int pnp_version(void);

/* ---- CSMRI data-fidelity gradient, fused with the variance-reduced update -----------------
 * Replaces CSMRI.grad_full (problems/CSMRI.py:76-81) and CSMRI.grad_stoch (:83-89), and the
 * update lines  z -= eta*lr_decay**i * v  of algorithms/pnp_gd.py:32-35, pnp_sgd.py:32-36,
 * pnp_svrg.py:53-57, pnp_sarah.py:72-75.
 *
 *   g   = Re(ifft2(sel o fft2(a - b) - Ysel))             (b, Y* optional)
 *   gs  = g * gscale
 *   v   = gs + vadd                                        (vadd optional)
 *   z_out = z_in - step * v                                (optional)
 * g_out / v_out receive gs / v when non-null. */
typedef struct {
    int H, W, batch;
    const float* a;               /* [batch][W][H] */
    const float* b;               /* optional: transform a - b (SVRG/SARAH difference, Y cancels) */
    float* S;                     /* scratch, batch*W*H floats (packed half spectrum, complex64) */
    const unsigned char* bits;    /* selection bits [batch][H/2][W], see pnp_csmri_sel_* */
    const float* Y1;              /* optional complex64 [batch][H/2][W]: (mask o Y)[kyp][kx]        */
    const float* Y2;              /*                                     conj (mask o Y)[-kyp][-kx]  */
    const float* Y1n;             /* optional complex64 [batch][W]: Nyquist row (mask o Y)[H/2][kx] */
    const float* Y2n;             /*                                conj (mask o Y)[H/2][-kx]       */
    float gscale;
    const float* gscale_ptr;      /* optional [batch], replaces gscale */
    float step;
    const float* step_ptr;        /* optional [batch], replaces step */
    float* g_out;
    const float* vadd;
    float* v_out;
    const float* z_in;
    float* z_out;
    int phases;                   /* 0 = all three passes; else bit0 lines r2c, bit1 columns+selection,
                                     bit2 lines c2r + epilogue (used to time the passes one by one).
                                     phases = 4 in the update form (vadd, z_in, z_out; no g_out / v_out) accepts
                                     S = NULL: a zero spectrum, z_out = z_in - step * vadd without the transform --
                                     the first inner iteration of a PnP-SVRG epoch (algorithms/pnp_svrg.py:53 with
                                     z == w), same bits as running the three passes on z - w = 0 */
    int clear_bits;               /* != 0: the column pass zeroes `bits` after using it (single-use minibatch
                                     selection; the next pnp_csmri_sel_* call then needs clear = 0) */
    /* Optional: build the minibatch selection INSIDE pass 1 (needs phases bit0), by every thread of the pass while
     * its first lines are in flight -- replaces a separate pnp_csmri_sel_* launch, which cannot share an SM with the
     * pass.  `bits` must then be writable and ALL ZERO on entry (clear_bits of the previous use leaves it so).
     * sel_count = 0: off.  Positions: sel_idx[img][*sel_cursor][0..sel_count) (k = ky*W + kx) when sel_idx is
     * non-null, else the keyed Feistel draw pnp_csmri_sel_sample makes from sel_support / sel_m0 / sel_seed /
     * *sel_counter (sel_count must not exceed sel_min_m0, the host-known lower bound of sel_m0[]). */
    int sel_count;
    const int* sel_idx;
    long long sel_idx_img_stride;
    const int* sel_cursor;
    const int* sel_support;
    const int* sel_m0;
    long long sel_support_img_stride;
    unsigned sel_seed;
    const int* sel_counter;
    int sel_min_m0;
    int flags;                    /* PNP_FLAG_* */
    /* Measurement shard (config 5): when row_hi > row_lo only the packed ky rows kyp in [row_lo, row_hi) of the half
     * spectrum are produced, masked and inverted -- kyp = min(ky, H - ky); row 0 also carries the Nyquist row -- i.e.
     * the result is the PARTIAL gradient of the measurements in that band (`bits` must select nothing outside it).
     * Pass 1 stores, pass 2 transforms and pass 3 loads only those rows.  0, 0: all rows. */
    int row_lo, row_hi;
} pnp_csmri_grad_args;
/* flags bit 0: launch the passes with programmatic dependent launch -- each kernel may begin (prologue, loads of data
 * older than its predecessor) while the previous kernel on the stream is finishing, and waits for it before it reads
 * its output.  Always safe; pays off when the stream is a chain of this library's passes (e.g. inside a CUDA graph). */
#define PNP_FLAG_CHAIN 1
int pnp_csmri_grad(const pnp_csmri_grad_args* args, void* stream);

/* ---- construction of a batch of CSMRI problems on the device (the step before the hot path) -------------------
 * Replaces, for sweeps that build hundreds of problems (script_diff_sampratio_set12.py:109-140 -> problems/CSMRI.py:12-41,
 * problems/problem.py:58-61), the NumPy constructor: mask ~ Bernoulli(p[img]), Y = mask o (fft2(X) + N(0, sigma)) with
 * sigma = sqrt(||mask o fft2(X)||_2 / 10^(snr/10) / H / W), Xinit = minmax(|ifft2(Y)|), support = flatnonzero(mask) --
 * with the library's own transforms and counter-based random numbers (a hash of seed, problem and k-space position; the
 * reference's sweeps are unseeded), straight into the layouts of pnp_csmri_grad / pnp_csmri_svrg_small.  No host
 * synchronisation: M0 and sigma stay on the device.
 *   x        [batch][W][H] float32 ground truth, line layout (in)
 *   p, snr   [batch] sampling probability / SNR in dB (device, in)
 *   bits_full [batch][H/2][W]; m0 [batch]; inv_m0 [batch] (optional); support [batch][support_img_stride] ascending, entries
 *            past m0 untouched (support_img_stride >= H*W guarantees room); Y1, Y2 [batch][H/2][W][2]; Y1n, Y2n
 *            [batch][W][2]; xinit [batch][W][H]; sigma [batch]
 *   work     scratch of pnp_csmri_build_batch_workspace(H, W, batch) bytes, 256-byte aligned */
typedef struct {
    int H, W, batch;
    unsigned seed;
    const float* x;
    const float* p;
    const float* snr;
    unsigned char* bits_full;
    int* m0;
    float* inv_m0;
    int* support;
    long long support_img_stride;
    float *Y1, *Y2, *Y1n, *Y2n;
    float* xinit;
    float* sigma;
    void* work;
} pnp_csmri_build_args;
long long pnp_csmri_build_batch_workspace(int H, int W, int batch);
int pnp_csmri_build_batch(const pnp_csmri_build_args* args, void* stream);

/* Selection bits from explicit k-space indices (k = ky*W + kx, the flat index into the
 * reference's (H, W) mask / minibatch arrays).  Replaces the dense 0/1 (H, W) minibatch array of
 * CSMRI.select_mb (problems/CSMRI.py:66-74) and the mask itself (:45).
 * idx: [batch][n_sets][count]; the set used is *cursor (0 when cursor is null).
 * `bits` is cleared first when clear != 0. */
int pnp_csmri_sel_from_indices(unsigned char* bits, int H, int W, int batch, const int* idx, int count,
                               long long idx_img_stride, const int* cursor, int clear, void* stream);

/* Device-drawn minibatch: `count` distinct entries of support[img][0..m0[img]) chosen by a keyed
 * Feistel permutation (seed, *counter).  Same role as np.random.choice(mask_locs, size,
 * replace=False) at problems/CSMRI.py:72 but not the same stream of numbers.
 * idx_out (optional, [batch][count]) receives the chosen k indices. */
int pnp_csmri_sel_sample(unsigned char* bits, int H, int W, int batch, const int* support, const int* m0,
                         long long support_img_stride, int count, unsigned seed, const int* counter,
                         int* idx_out, int clear, void* stream);

/* Device-drawn minibatch for problems whose measurements are a plain list (Problem.select_mb,
 * problems/problem.py:110-117): idx_out[0..count) = distinct positions in [0, n). */
int pnp_sample_indices(int* idx_out, int n, int count, unsigned seed, const int* counter, void* stream);

/* Host twin of the device sampler (same keyed Feistel permutation, bit-identical positions): fills
 * idx_out_host[0..count) on the CPU with `threads` worker threads.  Host plumbing for mb_source='host'
 * (the reference draws on the host too, problems/CSMRI.py:72); not a compute fallback. */
int pnp_sample_indices_host(int* idx_out_host, int n, int count, unsigned seed, unsigned counter, int img, int threads,
                            const int* support_host /* optional: idx_out = support_host[position] */);

/* Look-ahead queue of host draws for mb_source='host' (the reference draws one minibatch per iteration on the host,
 * problems/CSMRI.py:66-74 / problems/problem.py:110-117, inside the loop that waits for it; here `ahead` CPU worker
 * threads keep the next draws in flight while the GPU runs the current iteration).  Draw number c = 0, 1, 2, ... is
 * pnp_sample_indices_host(buffers[c % n_buffers], n, count, seed, c, 0, 1, support_host); the buffers belong to the
 * caller (pinned memory, count + extras ints each, n_buffers >= ahead + 2) and must outlive the handle.
 *   next   blocks until the next draw is complete and returns its slot; the caller must be done with a buffer
 *          n_buffers - ahead calls later (CPU-only use, no CUDA call is made)
 *   stage  next + write `extras` behind the indices + cudaMemcpyAsync(dst_dev <- buffer, count + n_extras ints) on
 *          `stream` + one event per slot, so a buffer is only handed to a new draw after the copy that read it
 *          has finished */
typedef struct pnp_host_draws pnp_host_draws;
int pnp_host_draws_create(pnp_host_draws** out, int n, int count, unsigned seed, const int* support_host,
                          int* const* buffers, int n_buffers, int ahead);
/* Draws of a queue created WITHOUT a host support list are ranks in [0, n).  With a device-resident support list set
 * here (the list of sampled k-space positions the device sampler uses; replaces the flatnonzero + choice of
 * problems/CSMRI.py:70-71), every _stage / _stage_many copy is followed, on the same stream, by an in-place
 * rank -> position gather on the device: the host decides WHICH measurements form the minibatch, the device resolves
 * where they sit -- the cache-missing gather (70 % of a 100k draw on one host core) leaves the host. */
int pnp_host_draws_set_device_support(pnp_host_draws* h, const int* support_dev);
int pnp_host_draws_next(pnp_host_draws* h, int* slot);
int pnp_host_draws_stage(pnp_host_draws* h, int* dst_dev, const int* extras, int n_extras, void* stream, int* slot);
/* n_draws consecutive draws (no extras) to dst_dev[j * dst_stride + 0 .. count), j = 0 .. n_draws - 1: like n_draws calls
 * of pnp_host_draws_stage, but draws that sit in consecutive staging buffers which are consecutive rows of ONE
 * allocation with row length dst_stride travel as one copy with one reuse event (the T2 minibatches of an SVRG epoch:
 * one or two copies instead of T2; at 256 x 256 the per-draw calls were 80 of the 140 us of host time per epoch). */
int pnp_host_draws_stage_many(pnp_host_draws* h, int* dst_dev, int n_draws, long long dst_stride, void* stream);
int pnp_host_draws_destroy(pnp_host_draws* h);

/* ---- Deblur + super-resolution gradient ------------------------------------------------------
 * Replaces Deblur.grad_full (problems/DeblurSR.py:126-132), Deblur.grad_stoch (:135-147) and the
 * same update lines as pnp_csmri_grad:
 *   x  = fft_blur(a - b, B)                      length-N circular convolution (:119-120)
 *   r  = S x - y on the selected measurements    (pylops Bilinear / Identity, :95-108)
 *   g  = fft_blur(S^T r, roll(flip(B), 1)) * gscale ;  v = g + vadd ;  z_out = z_in - step * v
 * sel = measurement ids [n_sets][count] (set *cursor), null = all M measurements. */
typedef struct {
    int H, W, batch;
    const float* a;
    const float* b;               /* optional: a - b, then y is ignored (use_y = 0) */
    float* S;                     /* scratch spectrum, H*W floats per image */
    float* blurred;               /* scratch image */
    float* up;                    /* scratch image (S^T r) */
    const float* Bf;              /* complex64 [H/2+1][W]: fft(B)[k1 + H*k2] * sqrt(N), row H/2 = Nyquist */
    const float* twn;             /* complex64 [W]: exp(-2 pi i j / (H*W)) */
    const float* y;               /* [M] measurements */
    const int* tl;                /* [M][2] top row / left column of the 4-tap footprint (null if identity) */
    const float* wts;             /* [M][2] weight of the lower row / right column */
    int identity;                 /* scale_percent == 100 */
    int M;
    const int* sel;
    int count;
    const int* cursor;
    int use_y;
    float gscale;
    float step;
    const float* step_ptr;
    float* g_out;
    const float* vadd;
    float* v_out;
    const float* z_in;
    float* z_out;
    int ntaps;                    /* > 0 (<= 16): the blur kernel has this many non-zero entries -> both convolutions */
    const int* tap_pos;           /* run as direct tap sums (HOST arrays: raveled positions p of B and the weights  */
    const float* tap_w;           /* B[p]*sqrt(N)); Bf / twn / S are then unused.  0: FFT path.                      */
} pnp_deblur_grad_args;
int pnp_deblur_grad(const pnp_deblur_grad_args* args, void* stream);

/* ---- Phase retrieval gradient (dense real Gaussian A, amplitude loss) ----------------------
 * Replaces PhaseRetrieval.grad_full (problems/PR.py:75-79) and grad_stoch (:81-87):
 *   t = A_sel z ;  r = ((|t| - y) / |t|) t  [ - the same at w when w != null ] ;  g = A_sel^T r * gscale
 * A: [M][n] float32 with columns in the line layout; rows = measurement ids (null = all). */
typedef struct {
    const float* A;
    long long n;
    int M;
    const float* z;
    const float* w;               /* optional second point (SVRG / SARAH difference; not linear) */
    const float* y;
    const int* rows;
    int count;
    const int* cursor;
    float* r;                     /* scratch, count floats */
    float gscale;
    float step;
    const float* step_ptr;
    float* g_out;
    const float* vadd;
    float* v_out;
    const float* z_in;
    float* z_out;
    /* Optional scratch of partial_chunks * n floats: the transposed product A_sel^T r is then split over row chunks as
     * well as columns (partial sums per chunk, added in order by a finishing kernel: deterministic) -- without it the
     * column pass has n / 1024 CTAs, 4 for a 64 x 64 image.  partial_chunks >= 1; 64 is what the Python layer passes. */
    float* partial;
    int partial_chunks;
} pnp_pr_grad_args;
int pnp_pr_grad(const pnp_pr_grad_args* args, void* stream);

/* Coded-diffraction phase retrieval with the intensity loss, ADDITIVE mode PhaseRetrieval(model='cdp') -- the
 * north star's "coded-diffraction |Ax|^2 - y Wirtinger gradient"; no counterpart in the reference (problems/PR.py
 * is the dense amplitude model above).  A_l x = fft2(d_l o x)/sqrt(N), d_l = i^codes[l], y = |A x|^2:
 *   g = gscale * Re( sum_l conj(d_l) o ifft2_unitary( m_l o (|A_l z|^2 - y_l) o A_l z ) )   [ - the same at w ]
 *   v = g + vadd ;  z_out = z_in - step * v
 * Images in the line layout; codes / y / mask: [L][W][H] in the same layout (int8 0..3 / float / byte).
 * sel_idx: `count` measurement ids m = l*H*W + ky*W + kx (set *cursor of [sets][count]); null = all M = L*H*W.
 * mask (L*H*W bytes, zero on entry, left zero) and S (L*H*W complex64) and acc (H*W floats) are scratch. */
typedef struct {
    int H, W, L;
    const signed char* codes;
    const float* y;
    const float* z;
    const float* w;               /* optional second point (SVRG / SARAH difference; not linear) */
    const int* sel_idx;
    int count;
    const int* cursor;
    unsigned char* mask;
    float* S;
    float* acc;
    float gscale;
    float step;
    const float* step_ptr;
    float* g_out;
    const float* vadd;
    float* v_out;
    const float* z_in;
    float* z_out;
    float* S2;                    /* optional second scratch like S: with w given, both points share the three passes'
                                     launches (same result as the two-sequence form) */
} pnp_cdp_grad_args;
int pnp_cdp_grad(const pnp_cdp_grad_args* args, void* stream);

/* ---- prox step ------------------------------------------------------------------------------
 * estimate_sigma(z0, multichannel=True, average_sigmas=True)  (algorithms/pnp_svrg.py:71 and
 * pnp_gd.py:49, pnp_sgd.py:50, pnp_saga.py:64, pnp_sarah.py:47,89).  ADDS the sum over columns of
 * the per-column estimates to sig_log[slot*batch + img] (double); the mean is that / W. */
int pnp_estimate_sigma(const float* z, int H, int W, int batch, double* sig_log, const int* slot,
                       void* stream);

/* TVDenoiser.denoise (denoisers/TV.py:21-26) = skimage denoise_wavelet(BayesShrink, db1, soft,
 * multichannel=True): per column multi-level Haar shrinkage.  sigma_est comes from
 * sig_log[slot*batch+img]/W when sig_log is non-null, else from sigma_est.  If it is > 0 the
 * threshold uses sigma_est*sigma_modifier, otherwise fallback_sigma (denoise_strength*decay**t).
 * When xrec and mse_log are non-null, sum((out - xrec)^2) is ADDED to mse_log[slot*batch+img]
 * (Problem.PSNR, problems/problem.py:33-35). */
int pnp_wavelet_denoise(const float* z_in, float* z_out, int H, int W, int batch, const double* sig_log,
                        float sigma_est, float sigma_modifier, float fallback_sigma, const float* xrec,
                        double* mse_log, const int* slot, void* stream);

/* estimate_sigma + TVDenoiser.denoise + PSNR in ONE cooperative launch (the iterate is read once, every
 * CTA keeps its lines in shared memory across a grid barrier).  Same results and log semantics as
 * pnp_estimate_sigma followed by pnp_wavelet_denoise with sig_log.  Returns PNP_ERR_UNSUPPORTED when the
 * image does not fit the SMs' shared memory (then use the two separate calls). */
int pnp_prox_wavelet_fused(const float* z_in, float* z_out, int H, int W, int batch, double* sig_log,
                           float sigma_modifier, float fallback_sigma, const float* xrec, double* mse_log,
                           const int* slot, void* stream);

/* The tail of a CSMRI inner iteration in ONE cooperative launch: pass 3 of pnp_csmri_grad (inverse line transforms
 * of the spectrum S that passes 1 and 2 left, phases = 3) + update z = z_in - step*(g*gscale + vadd) + sigma
 * estimate + wavelet BayesShrink + PSNR, with the iterate resident in shared memory in between (replaces
 * pnp_csmri_grad(phases = 4) followed by pnp_prox_wavelet_fused; same results).  batch 1.  Returns
 * PNP_ERR_UNSUPPORTED when the image lines do not fit the SMs' shared memory -- use the two calls then.
 * advance_counters (optional): the first n_advance ints are incremented at the end, as pnp_advance would (`slot`
 * may be one of them: it is read before).
 * S = NULL means a zero spectrum: z_out = Denoise(z_in - step * vadd) without the inverse transforms.  That is the
 * first inner iteration of every PnP-SVRG epoch (algorithms/pnp_svrg.py:53 with z == w: g_B(z) - g_B(w) is exactly 0,
 * so the three transform passes of that iteration are skipped; same bits as running them on zeros).
 * barrier_ws (optional): two zero-initialised 32-bit words (8-byte aligned: one 64-bit arrival counter) owned by the caller
 * and used with ONE image size only.  When given, the kernel is launched
 * normally and synchronises its CTAs (all co-resident: one per SM) with a software barrier on these words instead of a
 * cooperative launch -- the caller guarantees that no other kernel using such a barrier runs on the device at the same
 * time -- and `chain` != 0 adds programmatic dependent launch (see PNP_FLAG_CHAIN). */
int pnp_csmri_update_prox(const float* S, int H, int W, float gscale, float step, const float* step_ptr, const float* vadd,
                          const float* z_in, float* z_out, double* sig_log, float sigma_modifier, float fallback_sigma,
                          const float* xrec, double* mse_log, const int* slot, int* advance_counters, int n_advance,
                          unsigned* barrier_ws, int chain, void* stream);
/* The same launch, optionally followed -- inside the kernel, on the lines that are still in shared memory -- by the
 * FORWARD LINE PASS OF THE NEXT INNER ITERATION: S_out = packed half spectrum of (z_out - w), exactly what
 * pnp_csmri_grad(phases = 1, a = z_out, b = w) would store (the next iteration then starts at phases = 2), and by the
 * minibatch selection of that iteration (sel_* as in pnp_csmri_grad_args; the device sampler's draw counter is
 * *sel_counter + sel_counter_add because the counters advance inside this kernel).  Saves a kernel boundary (drain,
 * launch and cold start of a persistent pass: ~6 us at 2048 x 2048) and the re-read of the iterate.  S_out may be the
 * buffer S points to.  next = NULL: exactly pnp_csmri_update_prox.  PNP_ERR_UNSUPPORTED when a CTA holds more than 16
 * lines (one line per warp is what the kernel keeps resident). */
typedef struct {
    const float* w;
    float* S_out;
    unsigned char* bits;          /* selection bytes of the NEXT iteration (all zero on entry); used when sel_count > 0 */
    int sel_count;
    const int* sel_idx;           /* explicit positions [sel_count], or null: device sampler */
    const int* sel_support;
    const int* sel_m0;
    unsigned sel_seed;
    const int* sel_counter;
    int sel_counter_add;
    int sel_min_m0;
} pnp_csmri_next_pass;
int pnp_csmri_update_prox_next(const float* S, int H, int W, float gscale, float step, const float* step_ptr, const float* vadd,
                               const float* z_in, float* z_out, double* sig_log, float sigma_modifier, float fallback_sigma,
                               const float* xrec, double* mse_log, const int* slot, int* advance_counters, int n_advance,
                               unsigned* barrier_ws, int chain, const pnp_csmri_next_pass* next, void* stream);
/* 1 when pnp_csmri_update_prox handles H x W images on the current device, else 0 (host-side query, no launch) */
int pnp_csmri_update_prox_supported(int H, int W);

/* ---- whole PnP-SVRG runs of SMALL CSMRI images, one thread-block cluster per image ----------------------------
 * Replaces, for square images of 128 or 256 pixels a side, the whole loop of algorithms/pnp_svrg.py:26-95 in its
 * paper mode (v = (g_B(z) - g_B(w)) / B + mu, line 53) with the wavelet "TV" prox (denoisers/TV.py:21-26 after
 * estimate_sigma, pnp_svrg.py:71): `n_inner` inner iterations in ONE launch, a snapshot (mu = grad_full(z) *
 * snap_scale, w = z; pnp_svrg.py:32-35) before iteration 0 and then every T2 iterations, the step of epoch e being
 * step[img] * lr_decay**e.  The image, w and mu stay in the shared memory of a cluster of 8 CTAs for the whole run; the
 * transpositions of the 2-D transform go through distributed shared memory (csrc/small.cuh).  Same arithmetic as
 * pnp_csmri_grad + pnp_csmri_update_prox / pnp_prox_wavelet_fused.  `batch` images = `batch` clusters.
 * Minibatch of inner iteration t: idx[img * idx_img_stride + t * idx_iter_stride + 0..B) when idx is non-null, else the
 * keyed Feistel draw of pnp_csmri_sel_sample with counter  *draw_counter + t.  Logs: sig_log / mse_log[(*slot + t) *
 * batch + img] += the sum over lines of the sigma estimates / the squared error against xrec (mse_log and xrec
 * optional).  The counters are only read: advance them afterwards (pnp_advance_by).  fallback_sigma *
 * fallback_decay**t replaces a non-positive sigma estimate in iteration t.
 * Returns PNP_ERR_UNSUPPORTED for other sizes (use the per-pass entry points). */
typedef struct {
    int H, W, batch;
    float* z;                     /* [batch][W][H] iterate, updated in place */
    const float* xrec;            /* optional ground truth (PSNR log) */
    const float *Y1, *Y2, *Y1n, *Y2n;     /* as in pnp_csmri_grad_args */
    const unsigned char* bits_full;       /* selection bytes of the full mask [batch][H/2][W] */
    const int* support;           /* device sampler (idx == null): [batch][support_img_stride], m0[batch] */
    const int* m0;
    long long support_img_stride;
    const int* idx;               /* optional explicit minibatches */
    long long idx_img_stride, idx_iter_stride;
    const float* snap_scale_ptr;  /* optional [batch] (1 / M0), replaces snap_scale */
    float snap_scale;
    const float* step;            /* [batch] (step_img_stride = 1) or one value (0) */
    long long step_img_stride;
    double* sig_log;
    double* mse_log;
    const int* slot;              /* optional: first log slot (device int); null = 0 */
    const int* draw_counter;      /* optional: first draw counter (device int); null = 0 */
    int n_inner, T2, mini_batch_size;
    unsigned seed;
    float lr_decay, sigma_modifier, fallback_sigma, fallback_decay;
} pnp_csmri_svrg_small_args;
int pnp_csmri_svrg_small(const pnp_csmri_svrg_small_args* args, void* stream);
/* 1 when pnp_csmri_svrg_small handles H x W images, else 0 (host-side query, no launch) */
int pnp_csmri_svrg_small_supported(int H, int W);
/* how many images of that size the current device runs concurrently (co-resident clusters of 8 CTAs: 15 on a B200);
 * larger batches run in waves, and the per-pass entry points with the batch dimension then use the device better
 * (measured: 0.9 M image-iterations/s in waves of 15 against 1.1 M for two overlapped three-pass batches of 120) */
int pnp_csmri_svrg_small_capacity(int H, int W);

/* TV prox by Chambolle's dual projection, ADDITIVE mode (TVDenoiser(method='chambolle')): the north star's
 * "TV (Chambolle)" kernel; no counterpart in the reference, whose TVDenoiser is the wavelet shrink above
 * (denoisers/TV.py:24,26).  Arithmetic of scikit-image 0.18.2 denoise_tv_chambolle with a fixed n_iter >= 1
 * (no eps stop): z_out = z_in - div p after n_iter - 1 dual updates.  The weight is `weight` when > 0, else
 * sigma_est * sigma_modifier with sigma_est = sig_log[slot*batch+img]/W when that is > 0, else
 * fallback_weight; a weight <= 0 copies the input.  `work`: 4*H*W*batch floats of scratch (two ping-pong copies
 * of the dual field).  z_out must not alias z_in.  xrec / mse_log / slot as above. */
int pnp_tv_chambolle(const float* z_in, float* z_out, int H, int W, int batch, float weight, const double* sig_log,
                     float sigma_modifier, float fallback_weight, int n_iter, float* work, const float* xrec,
                     double* mse_log, const int* slot, void* stream);

/* NLMDenoiser.denoise (denoisers/NLM.py:22-27) = skimage denoise_nl_means(h = sigma = sigma_est *
 * sigma_modifier, fast_mode=False, patch_size, patch_distance) on a 2-D grey image.  Even patch sizes
 * are bumped to the next odd one as skimage does.  sigma_est <= 0 -> h = fallback_h, sigma = 0.
 * z_out must not alias z_in.  sig_log / xrec / mse_log / slot as in pnp_wavelet_denoise. */
int pnp_nlm_denoise(const float* z_in, float* z_out, int H, int W, int batch, int patch_size, int patch_distance,
                    const double* sig_log, float sigma_est, float sigma_modifier, float fallback_h,
                    const float* xrec, double* mse_log, const int* slot, void* stream);

/* ---- CNN denoisers (3x3 conv stacks, 64 features) ---------------------------------------------
 * Replaces self.model(x) + the wrapper arithmetic of RealSN_DnCNNDenoiser.denoise
 * (denoisers/RealSN_DnCNN.py:16-40; nets: DeepDenoisers/model/models.py:5-22, realSN_models.py:5-18,
 * SimpleCNN_models.py:6-56) and of MMODenoiser.denoise / apply_model (denoisers/MMODenoise.py:18-40,
 * 73-103,124-128).  Weights are packed by the host: layer 0 [9][64], middle layers [9][64 ci][64 co],
 * last layer [9][64]; tap = dl*3 + dp over (line, pixel-in-line) offsets of the transposed layout.
 * scale/shift = folded eval-mode BatchNorm or bias (null = 1 / 0); slope = activation (0 ReLU,
 * 0.01 LeakyReLU).  mode 0 = DnCNN wrapper (min/max normalise by *stats*, residual net),
 * mode 1 = MMO (clamp, net + input, clamp).  act0/act1: scratch, PH*PW*64 floats each; stats: 2 ints.
 * precision 0 = fp32 CUDA cores (exact-parity path); 1 = bf16 operands / fp32 accumulation on the tcgen05
 * tensor cores for the 64->64 layers: act0/act1 are then bf16 buffers of PH*(PW+1)*64 elements that the
 * caller ZEROES ONCE (one zero pad pixel per line is never written), and net->w_tc must be set;
 * 2 = error-compensated tensor-core mode ("bf16x3"): activations and weights as hi + lo bf16 pairs, three products per
 * tile into the fp32 accumulator -- fp32-path results to ~1e-5 (the reference runs these nets in fp32,
 * denoisers/RealSN_DnCNN.py:32-35): act0/act1 hold TWO such planes each (2*PH*(PW+1)*64 elements, zeroed once), and
 * net->w_tc_lo must be set as well. */
#define PNP_CNN_MAX_LAYERS 32
typedef struct {
    int n_layers;
    const float* w[PNP_CNN_MAX_LAYERS];
    const float* scale[PNP_CNN_MAX_LAYERS];
    const float* shift[PNP_CNN_MAX_LAYERS];
    float slope[PNP_CNN_MAX_LAYERS];
    float last_bias;
    int mode;
    float range, shift_in;
    const void* w_tc[PNP_CNN_MAX_LAYERS];   /* tensor-core path only (null otherwise), bf16, K = (dl, ci) contiguous:
                                               middle layers [192 rows (dp, co)][192] with scale[l] folded in,
                                               last layer [16 rows (dp, then zeros)][192] */
    const void* w_tc_lo[PNP_CNN_MAX_LAYERS]; /* precision 2 only: bf16(w - float(w_tc)), same layouts */
} pnp_cnn_net;
int pnp_cnn_forward(const pnp_cnn_net* net, const float* img, float* out, int PH, int PW, float* act0, float* act1,
                    int* stats, const float* xrec, double* mse_log, const int* slot, int precision, void* stream);

/* Problem.PSNR (problems/problem.py:33-35): ADDS sum((z - xrec)^2) to out[slot*batch+img]. */
int pnp_sq_err(const float* z, const float* xrec, long long n, int batch, double* out, const int* slot,
               void* stream);

/* ---- variance-reduction bookkeeping ---------------------------------------------------------
 * z_out = z_in - step * v   (algorithms/pnp_svrg.py:54-57 as committed: v = mu) */
int pnp_axpy(const float* z_in, const float* v, float* z_out, long long n, int batch, float step,
             const float* step_ptr, void* stream);
/* algorithms/pnp_saga.py:28 -- every table row starts as the same gradient */
int pnp_saga_init(const float* g0, float* table, float* tsum, long long n, int batch, int hist, void* stream);
/* algorithms/pnp_saga.py:45-50,72; g_prev is overwritten with g_new (prev_stoch = table[slot]) */
int pnp_saga_update(const float* g_new, float* g_prev, float* table, float* tsum, float* z,
                    long long n, int batch, int hist, const int* slot_idx, long long slot_img_stride,
                    const int* cursor, float step, const float* step_ptr, void* stream);
/* counters[0..n) += 1 on the device (log slot / minibatch cursor for CUDA-graph replay) */
int pnp_advance(int* counters, int n, void* stream);
/* counters[0..n) += delta (after a pnp_csmri_svrg_small run of delta iterations) */
int pnp_advance_by(int* counters, int n, int delta, void* stream);

/* counters[0..n) += 1 (n may be 0) and *x *= factor (step decay  eta*lr_decay**i  kept on the device) */
int pnp_advance_scale(int* counters, int n, float* x, float factor, void* stream);
/* dst[0..n) = src[0..n)  (w = copy(z), algorithms/pnp_svrg.py:35; capturable device copy) */
int pnp_copy_f32(float* dst, const float* src, long long n, void* stream);

/* debugging knobs, not part of the stable surface.  key 1: ablation bits of the tensor-core conv (1 no output
 * shift shuffles, 2 no epilogue arithmetic / stores, 4 load one dl block only, 32 never use the multi-layer launch);
 * key 2: print the phase times of the last pnp_csmri_update_prox (builds with -DPNP_PHASE_TIMING only). */
int pnp_debug_set(int key, int value);
/* key 1: copy the event trace of a -DPNP_TRACE build (16-byte records: u64 %globaltimer ns, i32 tag, i16 blockIdx.x,
 * i16 %smid) to out_host and reset it; returns the number of records (>= 0) or a negative status. */
int pnp_debug_read(int key, void* out_host, long long bytes);

/* ---- CUDA-graph helpers ---------------------------------------------------------------------
 * One inner iteration is launch-bound at 256x256 (about 2 MB of traffic); the host side captures
 * the iteration's launches once and replays the executable graph.  begin/end bracket the
 * launches on `stream`; `exec` is an opaque cudaGraphExec_t. */
int pnp_graph_begin(void* stream);
int pnp_graph_end(void* stream, void** exec_out);
int pnp_graph_launch(void* exec, void* stream);
int pnp_graph_destroy(void* exec);

#ifdef __cplusplus
}
#endif
#endif /* PNP_B200_H */
