/*
 * pileup_b200.h -- C ABI of the B200-native pile-up engine (libpileup_b200.so).
 *
 * The reference (open2c/coolpuppy 1.1.0) has no native plugin interface: the
 * seam this library replaces is the Python callable that
 * PileUpper.pileupsWithControl maps over view regions,
 *
 *     PileUpper.pileup_region(region1, region2, groupby, ...)      coolpup.py:1285-1358
 *       = CoordCreator.pos_stream  -> window dicts                  coolpup.py:598-746
 *       + PileUpper.get_data       -> region CSR                    coolpup.py:1024-1057
 *       + PileUpper._stream_snips  -> slice / mask / divide         coolpup.py:1059-1191
 *       + PileUpper.accumulate_stream + puputils._add_snip          coolpup.py:1236-1283, lib/puputils.py:12-41
 *
 * One call of pup_accumulate() does what one call of pileup_region() does for
 * the windows of one region: it adds every window into per-slot accumulators.
 * A "slot" is the host's dense id for (kind in {ROI, control}, group, flip);
 * the host maps slots back to the reference's {"ROI": {group: pup}, "control":
 * {...}} dictionary (coolpuppy_b200/coolpup.py).
 *
 * Conventions
 *  - Plain C types only.  Every data pointer may point to HOST memory (pageable
 *    or pinned) or to DEVICE memory on `device`; the library finds out with
 *    cudaPointerGetAttributes.  Caller keeps ownership of everything it passes.
 *  - All functions return PUP_OK (0) or a negative PUP_E_* code;
 *    pup_last_error() returns a thread-local description of the last failure.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Work is enqueued on it; calls that touch host buffers synchronise the
 *    stream before returning, calls on device buffers only enqueue.
 *  - Thread-safety: re-entrant; no global mutable state besides the
 *    thread-local error string.  A pup_region_t is immutable after creation
 *    and may be shared by concurrent pup_accumulate() calls on different
 *    streams of the same device.
 *  - There is no CPU fallback: without a CUDA device every compute entry point
 *    fails with PUP_E_NODEV.
 */
#ifndef PILEUP_B200_H
#define PILEUP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PUP_OK 0
#define PUP_E_ARG (-1)   /* bad argument (NULL, negative size, unsorted CSR, ...) */
#define PUP_E_CUDA (-2)  /* a CUDA runtime call or kernel failed */
#define PUP_E_OOM (-3)   /* device allocation failed */
#define PUP_E_NODEV (-4) /* no usable CUDA device */

/* flags: OOE / NODIAG are properties of a prepared region (pup_region_create), EXPCTRL / COVERAGE of a call
 * (pup_accumulate); pup_accumulate_region() takes the union. */
#define PUP_F_OOE 1u      /* divide every pixel by expected[|col-row|]         (coolpup.py:1154-1156) */
#define PUP_F_EXPCTRL 2u  /* also accumulate the bare expected block per window (coolpup.py:1135-1139, 1190-1191) */
#define PUP_F_COVERAGE 4u /* accumulate cov_start / cov_end                      (coolpup.py:1151-1153) */
#define PUP_F_NODIAG 8u   /* do NOT apply the signed diagonal mask (trans; unused by the cis path) */
#define PUP_F_LOCAL 32u   /* pup_accumulate_rescaled: symmetrise every snippet before the zoom (local pile-ups, coolpup.py:1215-1220) */
#define PUP_F_ASYNC 16u   /* pup_region_create / pup_accumulate with HOST input buffers: do not synchronise the
                             stream before returning; the caller keeps the (pinned) buffers alive and unchanged until
                             the stream has passed this call.  Lets uploads of region k+1 overlap the pile-up of k. */

typedef struct pup_region pup_region_t; /* opaque: a region matrix prepared in HBM */

int pup_abi_version(void);
const char* pup_last_error(void);
int pup_device_count(int* n_out);

/*
 * Upload and index one view region's contact matrix (replaces PileUpper.get_data +
 * cooler.Cooler.matrix(sparse=True).fetch(region).tocsr(), coolpup.py:1053-1057, and the
 * per-region bin vectors fetched at coolpup.py:1081-1098).
 *
 *   nb            number of bins of the region
 *   nnz           stored pixels of the SYMMETRIC-FILLED matrix (both triangles), < 2^31
 *   indptr[nb+1]  CSR row pointers, col[nnz] column ids sorted within each row, count[nnz] raw counts
 *   weight[nb]    balancing weights or NULL for raw counts; NaN marks a masked bin.
 *                 pixel value = (weight[row] * weight[col]) * count, as cooler computes it.
 *   expected[nb]  expected value by |col-row| or NULL; entries beyond the table must be NaN
 *   coverage[nb]  per-bin coverage (coverage_norm) or NULL
 *   ignore_diags  pixels with (col - row) < ignore_diags are masked (signed, coolpup.py:1141-1149)
 *   flags         PUP_F_OOE: pixel values are divided by expected[|col-row|]; PUP_F_NODIAG: no diagonal mask
 *
 * The normalisation (balancing, expected divide, diagonal mask, NaN -> "adds nothing") is applied here, once per
 * stored pixel, by a streaming kernel; the pile-up kernel then only gathers and adds.
 *
 * Besides the sparse strip layout a cis region (lower triangle masked) keeps a DENSE copy of the diagonals next to
 * the main diagonal, as far out as the matrix is dense: windows inside that band are piled up from it with register
 * tiles instead of the sparse scatter (same results up to fp64 summation order).  Memory budget: PUP_BAND_PCT percent
 * of the sparse pixel table (environment, default 100); the band ends where fewer than PUP_BAND_DENSITY_PCT percent
 * (default 20) of the cells are stored; PUP_BAND=0 builds none.
 */
int pup_region_create(int device, int32_t nb, int64_t nnz, const int32_t* indptr, const int32_t* col,
                      const int32_t* count, const double* weight, const double* expected,
                      const double* coverage, int ignore_diags, unsigned flags, void* stream,
                      pup_region_t** out);
/*
 * Same, from the UPPER triangle as cooler stores it (replaces the symmetric fill that cooler's
 * matrix(...).fetch() performs on the CPU, coolpup.py:1053-1057): row r of the region holds its stored pixels with
 * region-relative columns >= r, sorted; columns >= nb (pixels that leave the region: trans, or beyond a view arm)
 * are dropped.  With the signed diagonal mask and ignore_diags >= 0 (every cis pile-up of the reference) all pixels
 * below the diagonal are masked, so nothing is mirrored: the upload is indexed as it is, half the device work and
 * memory.  Otherwise (ignore_diags < 0 or PUP_F_NODIAG) the lower triangle is mirrored in on the device (stable
 * radix sort by column); device arrays are sized for the 2 * nnz_upper bound so no size is read back.
 * Both entry points drop the pixels the mask removes (col - row < ignore_diags) instead of storing zeros.
 */
int pup_region_create_upper(int device, int32_t nb, int64_t nnz_upper, const int32_t* indptr_upper,
                            const int32_t* col_upper, const int32_t* count_upper, const double* weight,
                            const double* expected, const double* coverage, int ignore_diags, unsigned flags,
                            void* stream, pup_region_t** out);
int pup_region_destroy(pup_region_t* region);
/*
 * Asynchronous host->device copy of a caller buffer (e.g. the window arrays of the region just created) through the
 * library's internal upload stream, the same FIFO the matrices of pup_region_create* travel on: the copy is served
 * in call order between the matrix uploads, and `stream` waits for it.  (A copy issued on any other stream can wait
 * on the copy engine until every queued matrix upload has been served.)  dst is DEVICE memory the caller owns and that no
 * pending work still uses (the upload stream does not wait for `stream`), src HOST memory (pinned for a truly
 * asynchronous copy) that must stay valid until `stream` has passed this call.
 */
int pup_upload(int device, void* dst, const void* src, int64_t bytes, void* stream);
/* bytes of HBM held by the region, and the algorithmic bytes of its pixels (8 per stored pixel) */
int64_t pup_region_device_bytes(const pup_region_t* region);

/*
 * Accumulator buffer: n_slots * pup_acc_stride(W) doubles, caller-owned (host or device), caller-zeroed;
 * pup_accumulate() only ever ADDS into it, so several regions (and, after an all-reduce, several GPUs)
 * can share one buffer -- the analogue of reduce(sum_pups) at coolpup.py:1511-1531.
 * The layout is linear in every field; decode it with pup_acc_export() AFTER all additions.
 */
int64_t pup_acc_stride(int W);

/*
 * Accumulate n_win windows of one region (replaces _stream_snips + accumulate_stream).
 *   r0[i], c0[i]   region-relative first row / first column bin of window i (W x W bins)
 *   slot[i]        accumulator slot in [0, n_slots)
 * Windows not fully inside [0, nb) are skipped and not counted (coolpup.py:1111-1114).
 *   flags          PUP_F_EXPCTRL and/or PUP_F_COVERAGE
 * n_valid_out (host pointer or NULL) receives the number of windows accumulated (forces a stream sync).
 */
int pup_accumulate(const pup_region_t* region, int64_t n_win, const int32_t* r0, const int32_t* c0,
                   const int32_t* slot, int W, int n_slots, unsigned flags, double* acc, void* stream,
                   int64_t* n_valid_out);

/* One-shot convenience: pup_region_create + pup_accumulate + pup_region_destroy. */
int pup_accumulate_region(int device, int32_t nb, int64_t nnz, const int32_t* indptr, const int32_t* col,
                          const int32_t* count, const double* weight, const double* expected,
                          const double* coverage, int64_t n_win, const int32_t* r0, const int32_t* c0,
                          const int32_t* slot, int W, int ignore_diags, int n_slots, unsigned flags,
                          double* acc, void* stream, int64_t* n_valid_out);

/*
 * Decode an accumulator buffer (host or device) into the reference's per-pile-up fields
 * (lib/puputils.py:12-38).  All outputs are HOST arrays, any of them may be NULL:
 *   sum[n_slots][W][W]      nansum of the snippets ("data")
 *   num[n_slots][W][W]      number of finite contributions per pixel ("num")
 *   n[n_slots]              number of windows ("n")
 *   cov_start/cov_end[n_slots][W]
 *   exp_sum/exp_num[n_slots][W][W]   sum / finite-count of the bare expected blocks (PUP_F_EXPCTRL)
 */
int pup_acc_export(const double* acc, int W, int n_slots, int device, void* stream, double* sum, int64_t* num,
                   int64_t* n, double* cov_start, double* cov_end, double* exp_sum, int64_t* exp_num);

/*
 * Rescaled pile-ups (PileUpper._rescale_snip, coolpup.py:1193-1234, for the windows of expand(.., rescale_flank),
 * 87-90, 108-114): window i is the h[i] x w[i] block at (r0[i], c0[i]) -- sizes differ per window -- built exactly as
 * _stream_snips builds a snippet (NaN for masked bins, masked diagonals, NaN expected; x / 0 = inf), optionally
 * symmetrised (PUP_F_LOCAL), then zoomed to rescale_size x rescale_size like cooltools.numutils.zoom_array
 * (scipy.ndimage.zoom(order=1) to the next multiple of rescale_size, block means; scipy's coordinate / weight
 * arithmetic is mirrored exactly), once for the snippet with NaN -> 0 and once for its NaN mask; output cells any NaN
 * touches are NaN (not summed, not counted), a snippet that is empty or all NaN counts as a block of zeros.
 * mode[i] = 1 (mode may be NULL): the snippet is the bare expected block E[|col - row|] instead of the matrix (the
 * "control" snippets of expected with ooe = False); needs a region created with an expected vector.
 * Accumulators: the layout of pup_acc_stride(rescale_size); sum and num (whole counts) per slot, n, and with
 * PUP_F_COVERAGE the zoomed coverage vectors.  Windows must be grouped by slot for speed (not for correctness).
 * r0 .. mode: host or device int32 arrays; acc: device memory.  Windows outside the region are dropped.
 */
int pup_accumulate_rescaled(const pup_region_t* region, int64_t n_win, const int32_t* r0, const int32_t* c0,
                            const int32_t* h, const int32_t* w, const int32_t* slot, const int32_t* mode, int rescale_size,
                            int n_slots, unsigned flags, double* acc, void* stream, int64_t* n_valid_out);

/*
 * store_stripes (coolpup.py:1164-1169): for every window the centre row `data[W/2, :]` ("horizontal") and the
 * reversed centre column `data[:, W/2][::-1]` ("vertical") of the snippet exactly as _stream_snips builds it
 * (NaN for masked bins / masked diagonals / NaN expected, x/0 = inf).  horizontal, vertical: [n_win][W] doubles,
 * host or device, in the order of the input windows; out-of-region windows give NaN rows.
 */
int pup_stripes(const pup_region_t* region, int64_t n_win, const int32_t* r0, const int32_t* c0, int W,
                double* horizontal, double* vertical, void* stream);

/*
 * Per-diagonal sums of one view region, the arithmetic of `cooltools expected-cis` whose table the reference takes
 * through --expected / expected_df (CLI.py:484-508, coolpup.py:861-918), from the region's upper triangle as cooler
 * stores it (same layout as pup_region_create_upper; columns >= nb are ignored).  Outputs, host or device, [nb] each:
 *   count_sum[d]     sum of raw counts over EVERY stored pixel of diagonal d, masked bins included (cooltools does
 *                    not mask raw counts; pinned by the reference's tests/data/CN.mm9.toy_expected.tsv)
 *   balanced_sum[d]  sum of (weight[row] * weight[col]) * count; NULL exactly when weight is NULL
 *   n_valid[d]       number of positions (i, i + d) whose two bins are valid (all nb - d without weights)
 * A bin is valid when its weight is not NaN.  expected = sum / n_valid (for count_sum too); masking the first
 * diagonals is the caller's.
 */
int pup_expected_cis(int device, int32_t nb, int64_t nnz_upper, const int32_t* indptr_upper, const int32_t* col_upper,
                     const int32_t* count_upper, const double* weight, double* count_sum, double* balanced_sum,
                     int64_t* n_valid, void* stream);

/*
 * Host-side window layout of one view region for bed features paired all-vs-all (replaces the pair loop of
 * CoordCreator.get_combinations and the row replication of _control_regions, coolpup.py:682-714, 387-453).  Plain
 * host code.  The m features of the region are sorted by position; pairs (k, k + i) are emitted by offset i = 1..m-1,
 * then k, and kept when mindist <= |center[k+i] - center[k]| <= maxdist.
 *   pup_pair_windows_count: per_offset[i] = kept pairs of offset i (per_offset[0] = 0); returns their total (-1 on bad
 *       arguments).  The caller draws the control shifts per offset with exactly these sizes (per_offset[i] * nctrl),
 *       which is what pins the reference's np.random stream.
 *   pup_pair_windows_fill: per offset block the ROI rows, then nctrl replicas shifted by dbin (block order, replica-
 *       major: dbin of block i starts after the draws of the previous blocks); outputs have
 *       total * (1 + nctrl) entries: st1 / st2 = first row / column bin (stbin + shift), kind (0 ROI, 1 control),
 *       idx1 / idx2 = feature indices k, l, distance = center[l] - center[k].
 */
int64_t pup_pair_windows_count(int32_t m, const double* center, double mindist, double maxdist, int64_t* per_offset);
/* the same, counting only the pairs whose first feature (the window's row anchor) k lies in [k_lo, k_hi) */
int64_t pup_pair_windows_count_range(int32_t m, const double* center, double mindist, double maxdist, int32_t k_lo,
                                     int32_t k_hi, int64_t* per_offset);
int pup_pair_windows_fill(int32_t m, const int64_t* stbin, const double* center, double mindist, double maxdist,
                          int32_t nctrl, const int64_t* dbin, int64_t* st1, int64_t* st2, int8_t* kind, int64_t* idx1,
                          int64_t* idx2, double* distance);

/*
 * Device-side window generation for bed features paired all-vs-all (no host loop, no window upload).
 *
 * pup_rng_t: numpy's legacy MT19937 state (np.random.get_state(): 624 key words + position) held on the device.  The
 * reference draws its control shifts from the GLOBAL np.random stream (coolpup.py:392-396, 442-445); a seeded run is
 * reproduced by loading that state, replaying the draws on the device and writing the final state back
 * (pup_rng_read -> np.random.set_state), so that host code after the pile-up continues the stream exactly where the
 * reference would.
 *
 * pup_control_shifts: for every segment s (one `_control_regions` call of the reference: a bedpe / local region, or
 * one pair offset of a bed region) seg_n[s] draws of `np.random.randint(minshift, maxshift, n)` followed by seg_n[s]
 * draws of `np.random.choice([-1, 1], n)`; dbin[draw] = np.round(shift * sign / resolution) (int32, device memory,
 * draws of all segments one after the other).  dbin == NULL only advances the generator (regions of other ranks).
 * seg_n is host memory; the call only enqueues work on `stream`.
 *
 * pup_pair_windows_device: the windows of CoordCreator.get_combinations (coolpup.py:682-714) + _control_regions
 * (387-453) for the m features of one view region (sorted like the reference sorts them), written in the
 * reference's emission order: pairs (k, k + i) by offset i then k, kept when mindist <= |center[k+i] - center[k]|
 * <= maxdist; per offset block the ROI rows, then nctrl replicas shifted by dbin (block order, replica-major:
 * draw = nctrl * base[i] + (rep - 1) * per_offset[i] + j).  Only the windows whose row anchor k lies in
 * [k_lo, k_hi) are written (a band of matrix rows: how a heavy region's windows are split over ranks), packed in
 * emission order; per_offset_part = pup_pair_windows_count_range(.., k_lo, k_hi, ..) (NULL: the whole region).
 *   stbin[m]        region-relative first bin of every feature's window; center[m] in bp; per_offset[m] from
 *                   pup_pair_windows_count (host memory; the others host or device)
 *   slot            = ((key * nk + kind) * nf + flip); key = key1[k] + key2[l] (swapped to key1[l] + key2[k] for a
 *                   flipped window when swap_on_flip: ignore_group_order, coolpup.py:131-144) + band_weight *
 *                   searchsorted(band_edges, center[l] - center[k], side="right") (bin_distance_intervals, 28-51);
 *                   NULL key arrays / band_edges contribute 0
 *   flip            flip_mode 0: never; 1: flipval[k] != 0 (flip_negative_strand: strand1 == "-"); 2: flipval[k] >
 *                   flipval[l] (flip_mark_intervals_func, 118-125)
 *   ident[m]        by-window (group_by_region, lib/puputils.py:218-223): every window is written twice, with keys
 *                   ident[k] and ident[l]; NULL otherwise
 *   first_seen      [n_keys] uint64, device, caller-initialised to all ones: atomicMin of
 *                   (kind << 62 | region_index << 40 | pos * targets + target) over the in-region windows of a key --
 *                   the order in which the reference's dictionaries first see a group; NULL: not recorded
 *   n_roi           [1] uint64, device: += number of in-region ROI windows (x targets); NULL: not counted
 * r0 / c0 / slot are device memory.  The call only enqueues work on `stream`; host arrays may be released on return.
 */
typedef struct pup_rng pup_rng_t;
int pup_rng_create(int device, const uint32_t* key624, int pos, void* stream, pup_rng_t** out);
int pup_rng_read(pup_rng_t* rng, uint32_t* key624, int* pos, void* stream);
int pup_rng_destroy(pup_rng_t* rng);
int pup_control_shifts(pup_rng_t* rng, int64_t n_segments, const int64_t* seg_n, int64_t minshift, int64_t maxshift,
                       double resolution, int32_t* dbin, void* stream);
int pup_pair_windows_device(int device, int32_t m, const int32_t* stbin, const double* center, double mindist,
                            double maxdist, int32_t nctrl, const int64_t* per_offset, const int32_t* dbin, int32_t nb,
                            int W, const int64_t* key1, const int64_t* key2, const double* band_edges, int32_t n_edges,
                            int64_t band_weight, int flip_mode, int swap_on_flip, const int32_t* flipval,
                            const int32_t* ident, int nk, int nf, int32_t k_lo, int32_t k_hi,
                            const int64_t* per_offset_part, int32_t region_index, int32_t* r0, int32_t* c0,
                            int32_t* slot, uint64_t* first_seen, uint64_t* n_roi, void* stream);

/* Statistics of the last pup_accumulate() on this thread (for bench.py): kernels launched by the call and
 * the exact algorithmic bytes of SURVEY.md section 8(d) -- filled only when n_valid_out was requested. */
int pup_last_launches(void);

/*
 * Optional device-side timing for benchmarks: when enabled (per host thread), every pup_accumulate() records
 * CUDA events on the caller's stream around its five phases: [0] window sort + chunk plan, [1] vector kernel,
 * [2] main (sparse) pile-up kernel, [3] dense-num kernel, [4] dense-band pile-up kernel.  pup_timing_read() waits for
 * the recorded events and returns the summed milliseconds and the number of spans per phase (arrays of 5); reset != 0
 * clears the records.
 */
int pup_timing_enable(int on);
int pup_timing_read(double* ms_by_phase, int* count_by_phase, int reset);

/*
 * Exact algorithmic bytes of a window list on a prepared region:
 *   sum over in-bounds windows of 16 + (W+1)*4 + 8*nnz_win (+16*W if balanced) (+16*W if coverage).
 * Computed on the device by a separate counting kernel (not part of the timed path).
 */
int pup_algorithmic_bytes(const pup_region_t* region, int64_t n_win, const int32_t* r0, const int32_t* c0, int W,
                          unsigned flags, void* stream, int64_t* bytes_out, int64_t* nnz_out);

#ifdef __cplusplus
}
#endif
#endif /* PILEUP_B200_H */
