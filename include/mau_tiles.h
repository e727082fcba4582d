/*
 * mau_tiles.h -- C ABI of the native tile reader (the input pipeline in front
 * of the hot path).
 *
 * Replaces, for one mini-batch at a time, what the reference does per sample on one Python
 * thread: `np.load(filepath)` of a `np.savez_compressed` archive and the four member reads
 * (reference src/dataset.py:54-59), the optional horizontal flip (src/dataset.py:139-141), the
 * `.float()` conversions (src/dataset.py:64-67), `torch.stack` of the batch and `pad_sequence`
 * of the temperature series (src/dataset.py:99-106).  The archive format is the one the
 * reference's writer produces (src/data/processing_10m/process.py:187): a ZIP (stored or
 * deflated members, ZIP64 records tolerated) holding `input.npy`, `target.npy`, `metadata.npy`
 * and `temperature_serie.npy` in NPY format 1.0 / 2.0 / 3.0.
 *
 * Host-only library (`libmau_tiles.so`, C++17 + zlib, no CUDA): it decodes straight into
 * caller-owned host memory (the Python side hands it pinned staging buffers and issues the
 * host->device copies itself).  Samples of a batch are decoded in parallel on a fixed pool of
 * worker threads; `submit` / `wait` let the caller keep several batches in flight.
 *
 * Conventions: every function returns 0 on success and a MAU_TILES_E_* code on failure;
 * mau_tiles_last_error() returns a thread-local NUL-terminated description.  Nothing throws
 * across the ABI.  All outputs are little-endian fp32, C-contiguous.
 */
#ifndef MAU_TILES_H_
#define MAU_TILES_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAU_TILES_OK          0
#define MAU_TILES_E_ARG       1  /* bad argument (null pointer, index out of range, ...)          */
#define MAU_TILES_E_IO        2  /* open / stat / mmap failed (np.load: FileNotFoundError / OSError) */
#define MAU_TILES_E_FORMAT    3  /* not a ZIP / NPY we can parse, CRC mismatch (zipfile.BadZipFile) */
#define MAU_TILES_E_MEMBER    4  /* a required member is missing (np.load: KeyError)               */
#define MAU_TILES_E_SHAPE     5  /* member shape differs from the batch shape (torch.stack: RuntimeError) */
#define MAU_TILES_E_CAPACITY  6  /* a temperature series is longer than series_stride              */
#define MAU_TILES_E_DTYPE     7  /* dtype / memory order we do not convert                         */

#define MAU_TILES_FLAG_NO_CRC 1  /* skip the CRC-32 check zipfile performs on every member read    */
#define MAU_TILES_FLAG_NICE    4 /* run the decode threads at niceness +10 (for hosts where the thread launching
                                    the GPU step would otherwise queue behind 6 ms decode tasks); measured
                                    neutral-to-worse on the 16-core B200 boxes, so off by default             */
#define MAU_TILES_FLAG_ZLIB   2  /* inflate through zlib instead of the reader's own decoder (A/B switch) */

typedef struct mau_tiles mau_tiles; /* opaque: a list of archive paths + the worker pool */

const char* mau_tiles_last_error(void);
int         mau_tiles_version(void);

/* The file list of one split (FuturePredictionDataset.__init__, src/dataset.py:35-36).  Paths are
 * copied.  threads <= 0 picks the number of online cores. */
int     mau_tiles_open(const char* const* paths, int64_t n, int threads, int flags, mau_tiles** out);
int     mau_tiles_close(mau_tiles* t);
int64_t mau_tiles_count(const mau_tiles* t);
int     mau_tiles_threads(const mau_tiles* t);

/* Shapes of sample `idx` read from the NPY headers only:
 * dims = { input C, H, W,  target C, H, W,  metadata length,  series length }.
 * (what data['input'].shape etc. would report, src/dataset.py:56-59) */
int mau_tiles_probe(mau_tiles* t, int64_t idx, int64_t dims[8]);

/* Lengths of the temperature series of n samples, read from the NPY headers on the worker pool (blocking).  A rank of a
 * data-parallel job pads its slice to the longest series of the GLOBAL batch (the reference's LSTM runs over the padding,
 * src/dataset.py:106, src/model.py:29-33) and needs the lengths of samples it does not decode. */
int mau_tiles_series_lengths(mau_tiles* t, const int64_t* idx, int64_t n, int64_t* out);

/* Decode n samples into batch-major buffers:
 *   input    [n, dims[0], dims[1], dims[2]]   (torch.stack(inputs),  src/dataset.py:99)
 *   target   [n, dims[3], dims[4], dims[5]]   (torch.stack(targets), src/dataset.py:101)
 *   metadata [n, dims[6]]                     (torch.stack(metadatas), src/dataset.py:100)
 *   series   [n, series_stride], zero-padded behind each sample's own length
 *            (pad_sequence(..., padding_value=0.0), src/dataset.py:106; the caller slices to the batch maximum)
 *   series_len[n]                             (temp_series_lengths, src/dataset.py:97)
 * dims[0..6] are the expected shapes (dims[7] is ignored); a sample that deviates fails the batch
 * with MAU_TILES_E_SHAPE.  hflip (may be NULL) holds one byte per sample: non-zero reverses the last
 * axis of input and target (RandomFlip, src/dataset.py:139-141).
 * Any of input / target / metadata / series may be NULL to skip that member. */
int mau_tiles_read_batch(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8],
                         float* input, float* target, float* metadata, float* series, int64_t series_stride,
                         int64_t* series_len);

/* Asynchronous form: returns a ticket > 0 (or -MAU_TILES_E_*); idx and hflip are copied, the output
 * buffers must stay valid until mau_tiles_wait(ticket) returns.  wait() returns the batch's status and
 * retires the ticket. */
int64_t mau_tiles_submit(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8],
                         float* input, float* target, float* metadata, float* series, int64_t series_stride,
                         int64_t* series_len);
/* The same with the input tiles delivered a second time in the engine's staged layout (include/mau_b200.h,
 * mau_plan_forward_staged): input_staged [n, dims[1], dims[2], staged_cs] bf16, NHWC, round-to-nearest-even, pad channels zero,
 * staged_cs a multiple of 8 and >= dims[0].  Converted by the thread that has just decoded the tile; a loader that ships this
 * buffer instead of `input` moves 2 * staged_cs instead of 4 * dims[0] bytes per pixel over PCIe (48 vs 92 bytes for the
 * 23-channel tiles).  `input` is still required (it is the decode target); input_staged == NULL is mau_tiles_submit. */
int64_t mau_tiles_submit_staged(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8],
                                float* input, uint16_t* input_staged, int64_t staged_cs, float* target, float* metadata,
                                float* series, int64_t series_stride, int64_t* series_len);
int     mau_tiles_wait(mau_tiles* t, int64_t ticket);
/* 1 if every sample of the batch has been decoded (wait() will not block), 0 if not, -MAU_TILES_E_ARG for an
 * unknown ticket.  Does not retire the ticket. */
int     mau_tiles_done(mau_tiles* t, int64_t ticket);

/* Rewrites archive `src` as `dst` with every member stored (method 0) -- what np.savez writes, still readable by
 * np.load and by the reference's loader -- so that later reads are a memcpy instead of an inflate (6.25 MB instead of
 * ~1.8 MB per 250 x 250 tile on disk).  All members are carried over (not only the four of a sample), CRCs are verified,
 * the file appears under its final name only when complete. */
int     mau_tiles_repack(const char* src_path, const char* dst_path);

/* The reader's DEFLATE decoder on its own (raw RFC 1951 stream `src` -> exactly dst_len bytes), decoded in two
 * calls split at output offset `split` (0 <= split <= dst_len) to exercise the resumable path; `src` must be
 * followed by 16 readable bytes.  Test / bench hook. */
int     mau_tiles_inflate(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len, size_t split);

/* The reader's CRC-32 (ZIP polynomial; zlib convention: pass 0 to start, the result to continue).  Test hook. */
uint32_t mau_tiles_crc32(uint32_t crc, const void* data, size_t n);

/* Bytes decoded (uncompressed NPY payload) and archive bytes consumed since open: loader bench. */
int     mau_tiles_stats(const mau_tiles* t, int64_t* payload_bytes, int64_t* archive_bytes, int64_t* samples);

#ifdef __cplusplus
}
#endif
#endif /* MAU_TILES_H_ */
