// pileup_b200.cu -- sm_100a kernels + C ABI of the pile-up engine (see include/pileup_b200.h).
//
// Hot path restated for the GPU (reference: coolpup.py:1059-1191 _stream_snips, 1236-1283 accumulate_stream,
// lib/puputils.py:12-41 _add_snip):
//
//   region prep (once per region):  every stored pixel is normalised exactly once
//       val = (w[row] * w[col]) * count / E[|col-row|];  val = 0 where the reference would produce NaN
//       (NaN weight, NaN expected, signed diagonal mask) because nansum() adds nothing there
//   storage: the matrix is cut into strips of R consecutive rows (R = 2 by default); inside a strip the pixels of
//       the R rows are merged and sorted by (col, row), so the part of a strip that a window needs is ONE
//       contiguous run R times longer than a CSR row run, read with full 128-byte lines
//   pile-up (k_pileup_main):
//   for every window (r0, c0, slot):                       # sorted by (slot, r0 mod R, r0, c0) on the device
//     for every strip G the window touches:                # one lane group (S = 4R lanes) owns R rows of the tile
//       start = bucket[(c0 >> lb) * ns + (r0 / R) + G]     # column-bucket-major strip pointer table: no binary search
//       stream the strip's (col, q, val) records from `start` while col < c0 + W   # 16-byte loads, 4 windows in flight
//       tile[G * R + q][col - c0] += val                   # fp64 shared-memory tile, race-free by row ownership
//   flush tile rows (tile row t = window row di + r0 mod R) with red.global.add.f64 when the slot changes
//   dense windows (k_pileup_dense): a window that lies inside the region's dense diagonal band is read cell by cell
//       from the band and added to per-thread REGISTER tiles -- no scatter at all (break-even ~18 % occupancy)
//
// `num` (count of finite contributions) is dense in the reference (W*W work per window).  Here it is
//   num = n_fast - rowbad[di] - colbad[dj] + xtile[di][dj]
// with a few masked-bin lookups per window (k_window_counts) incl. a sparse bad-row x bad-col correction; only windows that
// touch the masked diagonals / NaN expected values ("slow" windows) take the dense W*W path (k_num_slow).
#include "pileup_b200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
thread_local int g_launches = 0;

// optional per-phase device timing (bench.py): CUDA events recorded on the caller's stream
struct TimedSpan {
  int tag;  // 0 sort/plan, 1 vector kernel, 2 main (sparse) kernel, 3 dense-num kernel, 4 dense-band kernel
  cudaEvent_t a, b;
};
constexpr int N_TAGS = 5;
thread_local bool g_timing = false;
thread_local std::vector<TimedSpan> g_spans;

struct SpanGuard {
  bool on;
  TimedSpan sp;
  cudaStream_t st;
  SpanGuard(int tag, cudaStream_t s) : on(g_timing), st(s) {
    if (!on) return;
    sp.tag = tag;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, st);
  }
  void close() {
    if (!on) return;
    cudaEventRecord(sp.b, st);
    g_spans.push_back(sp);
    on = false;
  }
  ~SpanGuard() { close(); }
};

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess)
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  else
    snprintf(buf, sizeof buf, "%s", what);
  g_err = buf;
  return code;
}

#define CK(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? PUP_E_OOM : PUP_E_CUDA, #call, e_);        \
  } while (0)

#define LAUNCH_CHECK(name)                                                                     \
  do {                                                                                         \
    ++g_launches;                                                                              \
    cudaError_t e_ = cudaGetLastError();                                                       \
    if (e_ != cudaSuccess) return fail(PUP_E_CUDA, "launch " name, e_);                        \
  } while (0)

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}

// keep freed stream-ordered allocations in the device's default pool instead of returning them to the driver
// at every synchronisation (scratch buffers are re-used by the next call)
void retain_pool_memory(int dev) {
  static bool done[64] = {};
  if (dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    // never make one stream wait for another just to recycle a freed block: the upload stream of region k+1 must
    // not be chained behind the pile-up of region k (blocks whose free has already completed are still reused)
    if (env_int("PUP_POOL_DEPS", 0) == 0) {
      int off = 0;
      cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    }
  }
  cudaGetLastError();
  done[dev] = true;
}

// Host->device uploads run on an internal per-device copy stream so that they overlap whatever the caller's
// stream is still busy with (e.g. indexing the previous region); the caller's stream then waits on an event.
cudaStream_t copy_stream(int dev) {
  static cudaStream_t streams[64] = {};
  if (dev < 0 || dev >= 64) return nullptr;
  if (!streams[dev]) {
    if (cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError();
      streams[dev] = nullptr;
    }
  }
  return streams[dev];
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
    if (ok) retain_pool_memory(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// stream-ordered scratch allocations, released when the holder goes out of scope
struct Scratch {
  cudaStream_t stream;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
  cudaError_t alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, stream);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

// Zero-fill by a kernel, not cudaMemsetAsync: memsets are executed by a copy engine, where they queue behind the
// (hundreds of MB) host->device uploads of the following regions and stall the pile-up of the current one.
__global__ void k_zero(uint32_t* __restrict__ p, size_t n_words) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

cudaError_t zero_async(void* p, size_t bytes, cudaStream_t st) {
  const size_t n = (bytes + 3) / 4;  // all callers pass 4-byte-aligned buffers whose size is a multiple of 4
  if (n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  k_zero<<<grid, 256, 0, st>>>(reinterpret_cast<uint32_t*>(p), n);
  return cudaGetLastError();
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}


int ilog2_ceil(int64_t v) {
  int b = 0;
  while ((1ll << b) < v) ++b;
  return b;
}

constexpr int NT_MAX = 384;  // max threads per CTA of the main kernel (12 warps)
constexpr int SENT_PAD = 4;  // groups of S sentinel pixels at the start of pix[]
constexpr int VT = 512;      // max threads per CTA of the vector kernel
constexpr int VCH = 256;     // windows per CTA step of the vector kernel
constexpr int VU = 4;        // windows in flight per thread of the vector kernel

// one stored pixel as the main kernel reads it: 16 bytes, one ld.global.nc.v4 per lane
struct __align__(16) Pix {
  int col;
  int q;  // row inside its strip (row & (R - 1))
  double val;
};

// ------------------------------------------------------------------------------------------ accumulator layout
struct AccLayout {
  int W;
  int64_t w2, off_num, off_rb, off_cb, off_covs, off_cove, off_tsum, off_tnum, off_n, off_nfast, stride;
  __host__ __device__ explicit AccLayout(int W_) : W(W_) {
    w2 = (int64_t)W * W;
    off_num = w2;
    off_rb = 2 * w2;
    off_cb = off_rb + W;
    off_covs = off_cb + W;
    off_cove = off_covs + W;
    off_tsum = off_cove + W;
    off_tnum = off_tsum + 2 * W;
    off_n = off_tnum + 2 * W;
    off_nfast = off_n + 1;
    stride = off_n + 8;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------ region
struct pup_region {
  int device;
  int32_t nb;
  int64_t nnz;
  int lb;   // log2 of the column-bucket width
  int nbk;  // number of column buckets
  int R;    // rows per strip (power of two)
  int lr;   // log2 R
  int S;    // lanes per strip run in the main kernel; strips are aligned / padded to S pixels
  int32_t ns;  // number of strips = ceil(nb / R)
  int ignore_diags;
  unsigned flags;     // PUP_F_OOE | PUP_F_NODIAG folded into the pixel values
  Pix* pix;           // strip-major pixels: (col, q, normalised value), sorted by (col, q) inside a strip
  int32_t* prow;      // [ns+1] strip starts inside pix[]: multiples of S pixels, >= S sentinel pixels close a strip
  int32_t* bucket;    // [nbk][ns] first entry (rounded down to S) of strip s with col >= b << lb
  double* expected;   // [nb] or null
  double* coverage;   // [nb] or null
  uint8_t* bad;       // [nb] weight is NaN; null for raw counts
  int32_t* badpre;    // [nb+1] exclusive prefix count of masked bins; null for raw counts
  int32_t* badlist;   // sorted masked bins
  uint8_t* ebad;      // [nb] expected is NaN or 0
  int32_t* ebadpre;   // [nb+1] exclusive prefix of ebad
  double* band;       // dense diagonal band: band[row * band_stride + (col - row)] = normalised value (0 when
                      // unstored), for col - row in [0, *band_bw); null when the region has none
  int32_t* band_bw;   // device scalar: number of diagonals the band holds (chosen on the device from the density)
  int band_stride;    // row stride of band[] in doubles (the allocated width, >= *band_bw)
  cudaStream_t stream;
  int64_t bytes;
};

namespace {

// ------------------------------------------------------------------------------------------ prep kernels
// padded strip length (multiple of S pixels) -> scanned into prow[]
__global__ void k_padded_len(const int32_t* __restrict__ rs, const int32_t* __restrict__ re,
                             int32_t* __restrict__ plen, int nb, int ns, int lr, int S) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  // at least one complete group of S sentinel pixels follows the last stored pixel of every strip, so a lane that
  // steps past the end of its strip always reads a sentinel (col = INT_MAX) and the pile-up loop needs no
  // end-of-strip test
  if (s < ns) {
    const int r_lo = s << lr, r_hi = min(nb, r_lo + (1 << lr));
    int len = 0;
    for (int r = r_lo; r < r_hi; ++r) len += re[r] - rs[r];
    plen[s] = ((len + S - 1) / S + 1) * S;
  }
  if (s == ns) plen[s] = 0;
}

// first index in [lo, hi) of the sorted array `col` with col[idx] >= target (hi if none); the search gallops away
// from `hint` (neighbouring matrix rows have nearly the same profile, so the answer is a few entries from the hint)
__device__ __forceinline__ int lower_bound_hint(const int32_t* __restrict__ col, int lo, int hi, int hint, int target) {
  int pos = min(max(hint, lo), hi);
  int L, H;
  if (pos < hi && __ldg(&col[pos]) < target) {
    L = pos + 1;
    int step = 1;
    for (;;) {
      H = L + step;
      if (H >= hi) {
        H = hi;
        break;
      }
      if (__ldg(&col[H]) >= target) break;
      L = H + 1;
      step <<= 1;
    }
  } else {
    H = pos;
    int step = 1;
    for (;;) {
      L = H - step;
      if (L <= lo) {
        L = lo;
        break;
      }
      if (__ldg(&col[L]) < target) {
        L = L + 1;
        break;
      }
      H = L;
      step <<= 1;
    }
  }
  while (L < H) {
    const int mid = (L + H) >> 1;
    if (__ldg(&col[mid]) < target)
      L = mid + 1;
    else
      H = mid;
  }
  return L;
}

// One warp per strip: normalise every stored pixel once and write the 16-byte records of the strip's R rows merged
// in (col, row) order -- the position of a pixel is its index in its own row plus, for every other row of the
// strip, the number of that row's pixels that sort before it.  The strip is closed with sentinel pixels
// (col = INT_MAX, sorted last) up to the next multiple of S.
// Row r of the source matrix is col/cnt[rs[r] .. re[r]) (a CSR with rs = indptr, re = indptr + 1, or the in-region
// prefix of the rows of an uploaded upper triangle).
__global__ void k_prepare_pixels(const int32_t* __restrict__ rs, const int32_t* __restrict__ re,
                                 const int32_t* __restrict__ prow,
                                 const int32_t* __restrict__ col, const int32_t* __restrict__ cnt,
                                 const double* __restrict__ weight, const double* __restrict__ expected,
                                 Pix* __restrict__ pix, int nb, int ns, int lr, int ignore_diags, unsigned flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool ooe = (flags & PUP_F_OOE) && expected != nullptr;
  const bool nodiag = flags & PUP_F_NODIAG;
  if (warp == 0) {  // the sentinel block in front of the first strip
    for (int k = lane; k < prow[0]; k += 32) {
      Pix p;
      p.col = 0x7fffffff;
      p.q = 0;
      p.val = 0.0;
      pix[k] = p;
    }
  }
  for (int64_t s = warp; s < ns; s += nwarps) {
    const int r_lo = (int)s << lr, r_hi = min(nb, r_lo + (1 << lr));
    const int dst = prow[s], dend = prow[s + 1];
    int len = 0;
    for (int r = r_lo; r < r_hi; ++r) len += re[r] - rs[r];
    for (int r = r_lo; r < r_hi; ++r) {
      const int lo = rs[r], hi = re[r];
      double wr = 1.0;
      if (weight != nullptr) wr = weight[r];
      for (int i = lo + lane; i < hi; i += 32) {
        const int c = col[i];
        double v = (double)cnt[i];
        if (weight != nullptr) v = (wr * weight[c]) * v;  // cooler: bias[row] * bias[col] * count
        const int d = c - r;
        if (ooe) v = v / expected[d < 0 ? -d : d];
        if (!nodiag && d < ignore_diags) v = 0.0;  // signed diagonal mask (coolpup.py:1141-1149)
        if (v != v) v = 0.0;                       // NaN pixels (masked row / column bin, NaN expected) add nothing
        int rank = i - lo;
        for (int r2 = r_lo; r2 < r_hi; ++r2) {
          if (r2 == r) continue;
          const int lo2 = rs[r2], hi2 = re[r2];
          rank += lower_bound_hint(col, lo2, hi2, lo2 + (i - lo), c + (r2 < r ? 1 : 0)) - lo2;
        }
        Pix p;
        p.col = c;
        p.q = r - r_lo;
        p.val = v;
        pix[dst + rank] = p;
      }
    }
    for (int k = dst + len + lane; k < dend; k += 32) {
      Pix p;
      p.col = 0x7fffffff;
      p.q = 0;
      p.val = 0.0;
      pix[k] = p;
    }
  }
}

// Stored part of every source row: rs[r] = first entry with column >= r + min_d (pixels below that are masked by the
// signed diagonal rule, coolpup.py:1141-1149, and never contribute), re[r] = first entry with column >= nb (pixels
// of an uploaded upper triangle that leave the region).
__global__ void k_row_trim(const int32_t* __restrict__ indptr, const int32_t* __restrict__ col, int nb, int min_d,
                           int32_t* __restrict__ rs, int32_t* __restrict__ re) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb) return;
  const int lo0 = indptr[r], hi0 = indptr[r + 1];
  auto lower = [&](int target) {
    int lo = lo0, hi = hi0;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(&col[mid]) < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    return lo;
  };
  const int e = lower(nb);
  re[r] = e;
  rs[r] = min(lower(r + min_d), e);
}

// ---- symmetric fill on the device (cooler stores the upper triangle only; coolpup.py:1053-1057 relies on
// cooler's fetch to mirror it).  Input: upper CSR rows of the region, columns sorted, columns >= nb (pixels that
// leave the region, i.e. trans or beyond a view arm) are dropped.
__global__ void k_upper_counts(const int32_t* __restrict__ indptr_u, const int32_t* __restrict__ col_u, int nb,
                               int32_t* __restrict__ up_cnt, int32_t* __restrict__ lo_cnt,
                               int32_t* __restrict__ sort_key, int32_t* __restrict__ sort_val) {
  // one warp per row; sort_key/sort_val get the strictly-upper in-region entries' (col, entry index); entries that
  // are not mirrored get key nb (sorted last)
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nb; r += nwarps) {
    const int lo = indptr_u[r], hi = indptr_u[r + 1];
    int cnt = 0;
    for (int i = lo + lane; i < hi; i += 32) {
      const int c = col_u[i];
      const bool in = c >= 0 && c < nb;
      cnt += in;
      const bool mirror = in && c != (int)r;
      sort_key[i] = mirror ? c : nb;
      sort_val[i] = i;
      if (mirror) atomicAdd(&lo_cnt[c], 1);
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0) up_cnt[r] = cnt;
  }
}

__global__ void k_add_counts(const int32_t* a, const int32_t* b, int32_t* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
  if (i == n) out[i] = 0;
}

// row id of every upper entry (needed as the mirrored entry's column)
__global__ void k_expand_rows(const int32_t* __restrict__ indptr_u, int32_t* __restrict__ row_of, int nb) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nb; r += nwarps)
    for (int i = indptr_u[r] + lane; i < indptr_u[r + 1]; i += 32) row_of[i] = (int)r;
}

// upper part of every symmetric row: after the mirrored entries, same order as the input
__global__ void k_place_upper(const int32_t* __restrict__ indptr_u, const int32_t* __restrict__ col_u,
                              const int32_t* __restrict__ cnt_u, const int32_t* __restrict__ sym_indptr,
                              const int32_t* __restrict__ lo_cnt, int nb, int32_t* __restrict__ col_s,
                              int32_t* __restrict__ cnt_s) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nb; r += nwarps) {
    const int lo = indptr_u[r], hi = indptr_u[r + 1];
    const int dst = sym_indptr[r] + lo_cnt[r];
    for (int i = lo + lane; i < hi; i += 32) {
      const int c = col_u[i];
      if (c >= 0 && c < nb) {  // in-region columns form a prefix of the sorted row
        col_s[dst + (i - lo)] = c;
        cnt_s[dst + (i - lo)] = cnt_u[i];
      }
    }
  }
}

// mirrored part: entries stably sorted by column arrive grouped by target row, source rows ascending
__global__ void k_place_lower(const int32_t* __restrict__ sorted_key, const int32_t* __restrict__ sorted_val,
                              const int32_t* __restrict__ row_of, const int32_t* __restrict__ cnt_u,
                              const int32_t* __restrict__ sym_indptr, const int32_t* __restrict__ lo_start,
                              int nb, int32_t* __restrict__ col_s, int32_t* __restrict__ cnt_s) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= lo_start[nb]) return;  // lo_start[nb] = number of mirrored entries (they sort first)
  const int c = sorted_key[j];
  const int src = sorted_val[j];
  const int dst = sym_indptr[c] + (int)(j - lo_start[c]);
  col_s[dst] = row_of[src];
  cnt_s[dst] = cnt_u[src];
}

// bucket[b * ns + s] = position in pix[] (rounded down to a multiple of S pixels inside the strip) of the first
// pixel of strip s whose column is >= (b << lb).  One streaming pass over the strip-major pixels: wherever the
// column bucket changes between neighbouring pixels, the thread of the later pixel writes the entries of all
// buckets in between; the sentinel that closes a strip (col = INT_MAX) fills the strip's remaining buckets.
__global__ void k_build_buckets(const Pix* __restrict__ pix, const int32_t* __restrict__ prow,
                                int32_t* __restrict__ bucket, int ns, int nbk, int lb, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < ns; s += nwarps) {
    const int lo = prow[s], hi = prow[s + 1];
    for (int base = lo; base < hi; base += 32) {
      const int i = base + lane;
      if (i >= hi) break;
      const int c = __ldg(&pix[i].col);
      const int bc = min(c >> lb, nbk - 1);
      const int bp = (i == lo) ? -1 : min(__ldg(&pix[i - 1].col) >> lb, nbk - 1);
      for (int b = bp + 1; b <= bc; ++b) bucket[(int64_t)b * ns + s] = i & ~(S - 1);
      if (c == 0x7fffffff) break;  // first sentinel reached: all buckets of this strip are written
    }
  }
}

__global__ void k_badlist(const uint8_t* __restrict__ bad, const int32_t* __restrict__ badpre,
                          int32_t* __restrict__ badlist, int nb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb && bad[i]) badlist[badpre[i]] = i;
}

__global__ void k_masks(const double* __restrict__ weight, const double* __restrict__ expected, uint8_t* bad,
                        uint8_t* ebad, int32_t* ebad32, int32_t* bad32, int nb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nb) return;
  if (i == nb) {
    ebad32[i] = 0;
    if (bad32 != nullptr) bad32[i] = 0;
    return;
  }
  if (bad != nullptr) {
    const int b = isnan(weight[i]) ? 1 : 0;
    bad[i] = (uint8_t)b;
    bad32[i] = b;
  }
  uint8_t eb = 0;
  if (expected != nullptr) {
    double e = expected[i];
    eb = (isnan(e) || e == 0.0) ? 1 : 0;
  }
  ebad[i] = eb;
  ebad32[i] = eb;
}

// ------------------------------------------------------------------------------------------ window keys
// key = [invalid:1][slot][r0 mod R : lr][r0:pb][c0:pb]; out-of-region windows get all ones and sort last.
// (slot, r0 mod R) is the "extended slot": all windows of one extended slot map matrix rows to tile rows the same
// way (tile row = window row + r0 mod R), so the main kernel treats a change of either like a change of slot.
// With a dense diagonal band (band_bw != null) the windows that lie entirely inside it form a second class, keyed as
// slots [n_slots, 2 n_slots): they sort behind all other windows and are piled up by k_pileup_dense.
__global__ void k_window_keys(const int32_t* __restrict__ r0, const int32_t* __restrict__ c0,
                              const int32_t* __restrict__ slot, uint64_t* __restrict__ keys, int64_t n, int nb, int W,
                              int n_slots, int pb, int lr, const int32_t* __restrict__ band_bw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = r0[i], c = c0[i], s = slot[i];
  bool ok = r >= 0 && c >= 0 && r + W <= nb && c + W <= nb && s >= 0 && s < n_slots;
  if (band_bw != nullptr) {
    const int bw = __ldg(band_bw), D0 = c - r;
    if (D0 - (W - 1) >= 0 && D0 + (W - 1) < bw) s += n_slots;
  }
  const uint64_t es = ((uint64_t)s << lr) | (uint64_t)(r & ((1 << lr) - 1));
  keys[i] = ok ? ((es << (2 * pb)) | ((uint64_t)r << pb) | (uint64_t)c) : ~0ull;
}

// sorted keys -> (r0, c0) records, so that the main kernel does no 64-bit key arithmetic
__global__ void k_decode_windows(const uint64_t* __restrict__ keys, int2* __restrict__ win, int64_t n, int pb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  const uint64_t m = (1ull << pb) - 1;
  win[i] = make_int2((int)((k >> pb) & m), (int)(k & m));
}

// slot_start[s] = first sorted window of extended slot s (s = n_slots: number of valid windows);
// nchunks[s] = ceil(count / ch).  n_slots counts EXTENDED slots here.
__global__ void k_slot_bounds(const uint64_t* __restrict__ keys, int n, int n_slots, int pb, int ch,
                              int32_t* __restrict__ slot_start, int32_t* __restrict__ nchunks) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_slots) return;
  auto lower = [&](int sl) {
    uint64_t target = (sl >= n_slots) ? (1ull << 63) : ((uint64_t)sl << (2 * pb));
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (keys[mid] < target)
        lo = mid + 1;
      else
        hi = mid;
    }
    return lo;
  };
  int a = lower(s);
  slot_start[s] = a;
  if (s < n_slots) {
    int b = lower(s + 1);
    nchunks[s] = (b - a + ch - 1) / ch;
  } else {
    nchunks[s] = 0;
  }
}

// ------------------------------------------------------------------------------------------ shared device helpers
struct WinCtx {
  int nb, W, pb, lr, ignore_diags;
  unsigned flags;
  const int32_t* ebadpre;
  int n_slots;  // accumulator slots; keyed slots [n_slots, 2 n_slots) are the dense-band class of the same slots
};

// sorted key -> accumulator slot (the r0 mod R bits of the extended slot are dropped), r0, c0
__device__ __forceinline__ void decode_key(uint64_t k, const WinCtx& c, int& slot, int& r0, int& c0) {
  uint64_t m = (1ull << c.pb) - 1;
  c0 = (int)(k & m);
  r0 = (int)((k >> c.pb) & m);
  slot = (int)(k >> (2 * c.pb + c.lr));
  if (slot >= c.n_slots) slot -= c.n_slots;
}

// A window is "slow" when some pixel is masked by the signed diagonal rule or by a NaN/zero expected value:
// its `num` contribution is then evaluated densely.  Fast windows contribute through rb/cb/n_fast only.
__device__ __forceinline__ bool window_is_slow(const WinCtx& c, int r0, int c0) {
  int D0 = c0 - r0;
  int dmin = D0 - (c.W - 1), dmax = D0 + (c.W - 1);
  if (!(c.flags & PUP_F_NODIAG) && dmin < c.ignore_diags) return true;
  if (c.flags & PUP_F_OOE) {
    int a, b;
    if (dmin >= 0) {
      a = dmin;
      b = dmax;
    } else if (dmax <= 0) {
      a = -dmax;
      b = -dmin;
    } else {
      a = 0;
      b = max(-dmin, dmax);
    }
    b = min(b, c.nb - 1);
    a = min(a, b);
    return (__ldg(&c.ebadpre[b + 1]) - __ldg(&c.ebadpre[a])) > 0;
  }
  return false;
}

// work item -> (extended slot, window range) through the per-extended-slot chunk table
struct ChunkTable {
  const int32_t* slot_start;   // [n_slots+1]
  const int32_t* chunk_start;  // [n_slots+1] exclusive scan of chunks per extended slot
  int n_slots;                 // number of EXTENDED slots (accumulator slots << lr)
  int ch;  // windows per chunk
};

__device__ __forceinline__ void locate_chunk(const ChunkTable& t, int chunk, int& slot, int& lo, int& hi) {
  int a = 0, b = t.n_slots;  // last s with chunk_start[s] <= chunk
  while (a < b) {
    int mid = (a + b + 1) >> 1;
    if (__ldg(&t.chunk_start[mid]) <= chunk)
      a = mid;
    else
      b = mid - 1;
  }
  slot = a;
  lo = __ldg(&t.slot_start[a]) + (chunk - __ldg(&t.chunk_start[a])) * t.ch;
  hi = min(lo + t.ch, __ldg(&t.slot_start[a + 1]));
}

// ------------------------------------------------------------------------------------------ window counts
// Everything `num` needs from a fast window besides the pixels: which of its rows / columns are masked bins.
// Masked bins are sparse, so instead of scanning 2W flags per window one thread per window looks its masked rows
// and columns up in the region's sorted masked-bin list (two prefix-count reads per side) and adds them to
// privatised int32 tiles: rb[di], cb[dj] and the masked-row x masked-column correction xb[di][dj].
// Layout of one private copy: [n_slots][W*W + 2*W + 2] ints = xb | rb | cb | n_slow | pad.
struct CountParams {
  WinCtx ctx;
  const uint64_t* keys;
  const int32_t* slot_start;  // [n_eslots+1]
  int n_slots;                // accumulator slots
  int n_eslots;               // extended slots = n_slots << lr
  const int32_t* badpre;     // [nb+1] exclusive prefix count of masked bins; null for raw counts
  const int32_t* badlist;    // sorted masked bins
  int* counts;               // [copies][n_slots][cstride]
  int copies;
  int* n_slow;               // number of slow windows of this call
};

__device__ __forceinline__ int64_t count_stride(int W) { return (int64_t)W * W + 2 * W + 2; }

__global__ void __launch_bounds__(256) k_window_counts(const CountParams p) {
  const int n = __ldg(&p.slot_start[p.n_eslots]);
  const int W = p.ctx.W;
  const int64_t cs = count_stride(W);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool slow = false;
  if (i < n) {
    int slot, r0, c0;
    decode_key(__ldg(&p.keys[i]), p.ctx, slot, r0, c0);
    slow = window_is_slow(p.ctx, r0, c0);
    int* base = p.counts + ((int64_t)(blockIdx.x % p.copies) * p.n_slots + slot) * cs;
    if (slow) {
      atomicAdd(base + (int64_t)W * W + 2 * W, 1);
    } else if (p.badpre != nullptr) {
      const int a0 = __ldg(&p.badpre[r0]), a1 = __ldg(&p.badpre[r0 + W]);
      const int b0 = __ldg(&p.badpre[c0]), b1 = __ldg(&p.badpre[c0 + W]);
      int* rb = base + (int64_t)W * W;
      int* cb = rb + W;
      for (int k = a0; k < a1; ++k) atomicAdd(rb + (__ldg(&p.badlist[k]) - r0), 1);
      for (int k = b0; k < b1; ++k) atomicAdd(cb + (__ldg(&p.badlist[k]) - c0), 1);
      for (int k = a0; k < a1; ++k) {
        const int di = __ldg(&p.badlist[k]) - r0;
        for (int l = b0; l < b1; ++l) atomicAdd(base + (int64_t)di * W + (__ldg(&p.badlist[l]) - c0), 1);
      }
    }
  }
  if (__syncthreads_or(slow) && threadIdx.x == 0) atomicAdd(p.n_slow, 1);  // "this call has slow windows"
}

// The same counts through a shared-memory tile: a CTA walks one contiguous span of the sorted windows (one slot at a
// time, windows being sorted by slot), adds with native shared-memory integer atomics and flushes the non-zero entries
// of the tile to its private global copy when the slot changes.  ~12 global atomics per window become ~12 shared ones
// (configs[3]: the count phase 1.38 -> 1.16 ms per step).  Windows of another slot than the CTA's current one (span boundaries) go to
// the global copy directly.
__global__ void __launch_bounds__(256) k_window_counts_smem(const CountParams p) {
  extern __shared__ int ctile[];
  __shared__ int s_first;
  const int n = __ldg(&p.slot_start[p.n_eslots]);
  const int W = p.ctx.W;
  const int64_t cs = count_stride(W);
  const int per = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  for (int i = threadIdx.x; i < (int)cs; i += blockDim.x) ctile[i] = 0;
  int* copy = p.counts + (int64_t)(blockIdx.x % p.copies) * p.n_slots * cs;
  int cur = -1;
  bool any_slow = false;
  auto flush = [&]() {
    if (cur < 0) return;
    int* dst = copy + (int64_t)cur * cs;
    for (int i = threadIdx.x; i < (int)cs; i += blockDim.x) {
      const int v = ctile[i];
      if (v != 0) {
        atomicAdd(dst + i, v);
        ctile[i] = 0;
      }
    }
  };
  __syncthreads();
  for (int base = lo; base < hi; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int slot = -1, r0 = 0, c0 = 0;
    if (i < hi) decode_key(__ldg(&p.keys[i]), p.ctx, slot, r0, c0);
    if (threadIdx.x == 0) s_first = slot;
    __syncthreads();
    if (s_first != cur) {  // uniform: every thread reads the same s_first
      flush();
      cur = s_first;
      __syncthreads();
    }
    if (i < hi) {
      const bool mine = slot == cur;
      int* base_t = mine ? ctile : copy + (int64_t)slot * cs;
      const bool slow = window_is_slow(p.ctx, r0, c0);
      any_slow |= slow;
      if (slow) {
        atomicAdd(base_t + (int64_t)W * W + 2 * W, 1);
      } else if (p.badpre != nullptr) {
        const int a0 = __ldg(&p.badpre[r0]), a1 = __ldg(&p.badpre[r0 + W]);
        const int b0 = __ldg(&p.badpre[c0]), b1 = __ldg(&p.badpre[c0 + W]);
        int* rb = base_t + (int64_t)W * W;
        int* cb = rb + W;
        for (int k = a0; k < a1; ++k) atomicAdd(rb + (__ldg(&p.badlist[k]) - r0), 1);
        for (int k = b0; k < b1; ++k) atomicAdd(cb + (__ldg(&p.badlist[k]) - c0), 1);
        for (int k = a0; k < a1; ++k) {
          const int di = __ldg(&p.badlist[k]) - r0;
          for (int l = b0; l < b1; ++l) atomicAdd(base_t + (int64_t)di * W + (__ldg(&p.badlist[l]) - c0), 1);
        }
      }
    }
    __syncthreads();  // the tile may be flushed at the top of the next round
  }
  flush();
  if (__syncthreads_or(any_slow) && threadIdx.x == 0) atomicAdd(p.n_slow, 1);  // "this call has slow windows"
}

// acc += private copies; n comes from the slot boundaries, n_fast = n - n_slow
__global__ void k_counts_reduce(const int* __restrict__ counts, int copies, int n_slots, int W, int lr, int n_cls,
                                const int32_t* __restrict__ slot_start, double* __restrict__ acc) {
  const AccLayout L(W);
  const int64_t cs = count_stride(W);
  const int64_t total = (int64_t)n_slots * cs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int sum = 0;
    for (int c = 0; c < copies; ++c) sum += counts[(int64_t)c * total + i];
    const int64_t s = i / cs, j = i - s * cs;
    double* a = acc + s * L.stride;
    if (j < L.w2) {
      if (sum) atomicAdd(a + L.off_num + j, (double)sum);
    } else if (j < L.w2 + W) {
      if (sum) atomicAdd(a + L.off_rb + (j - L.w2), (double)sum);
    } else if (j < L.w2 + 2 * W) {
      if (sum) atomicAdd(a + L.off_cb + (j - L.w2 - W), (double)sum);
    } else if (j == L.w2 + 2 * W) {
      int nwin = 0;  // over the slot's extended slots, in both window classes
      for (int c = 0; c < n_cls; ++c)
        nwin += slot_start[(c * n_slots + s + 1) << lr] - slot_start[(c * n_slots + s) << lr];
      if (nwin) {
        atomicAdd(a + L.off_n, (double)nwin);
        atomicAdd(a + L.off_nfast, (double)(nwin - sum));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ vector kernel
// Per-window O(W) fp64 quantities, only launched when requested: coverage sums (PUP_F_COVERAGE) and the Toeplitz
// sums of the bare expected block (PUP_F_EXPCTRL).  One CTA walks VCH consecutive sorted windows, VU at a time;
// thread t owns vector index t.
struct VecParams {
  WinCtx ctx;
  const uint64_t* keys;
  const int32_t* slot_start;  // [n_eslots+1]
  int n_eslots;
  const double* expected;    // for EXPCTRL
  const double* coverage;    // for COVERAGE
  double* acc;
};

__global__ void k_vector(const VecParams p) {
  const int n = __ldg(&p.slot_start[p.n_eslots]);
  const int W = p.ctx.W;
  const AccLayout L(W);
  const int t = threadIdx.x;
  const bool cov = (p.ctx.flags & PUP_F_COVERAGE) && p.coverage != nullptr;
  const bool ectl = (p.ctx.flags & PUP_F_EXPCTRL) && p.expected != nullptr;
  for (int base = blockIdx.x * VCH; base < n; base += gridDim.x * VCH) {
    const int end = min(base + VCH, n);
    int cur = -1;
    double cs = 0, ce = 0, ts = 0, tn = 0;
    auto flush = [&]() {
      if (cur < 0) return;
      double* a = p.acc + (int64_t)cur * L.stride;
      if (t < W) {
        if (cs != 0) atomicAdd(a + L.off_covs + t, cs);
        if (ce != 0) atomicAdd(a + L.off_cove + t, ce);
      }
      if (t < 2 * W - 1) {
        if (ts != 0) atomicAdd(a + L.off_tsum + t, ts);
        if (tn != 0) atomicAdd(a + L.off_tnum + t, tn);
      }
      cs = ce = ts = tn = 0;
    };
    for (int w = base; w < end; w += VU) {
      int slot[VU];
      double ca[VU], cbv[VU], ev[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) {
        slot[u] = -1;
        ca[u] = cbv[u] = 0.0;
        ev[u] = 0.0;
        if (w + u < end) {
          int r0, c0;
          decode_key(__ldg(&p.keys[w + u]), p.ctx, slot[u], r0, c0);
          if (cov && t < W) {
            ca[u] = __ldg(&p.coverage[r0 + t]);
            cbv[u] = __ldg(&p.coverage[c0 + t]);
          }
          if (ectl && t < 2 * W - 1) {
            int d = c0 - r0 + t - (W - 1);
            ev[u] = __ldg(&p.expected[d < 0 ? -d : d]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < VU; ++u) {
        if (slot[u] < 0) continue;
        if (slot[u] != cur) {
          flush();
          cur = slot[u];
        }
        if (cov && t < W) {
          if (!isnan(ca[u])) cs += ca[u];
          if (!isnan(cbv[u])) ce += cbv[u];
        }
        if (ectl && t < 2 * W - 1) {
          if (!isnan(ev[u])) ts += ev[u];
          if (isfinite(ev[u])) tn += 1;
        }
      }
    }
    flush();
  }
}

// ------------------------------------------------------------------------------------------ main kernel
struct MainParams {
  int W, ns, lb;
  const Pix* pix;
  const int32_t* bucket;
  const int2* win;  // sorted (r0, c0)
  ChunkTable chunks;  // over extended slots
  int Gb;        // lane groups (= strips of a window) per band
  int n_groups;  // strips a window can touch: ceil((W + R - 1) / R)
  int n_bands;
  int TW;        // tile row stride in doubles (>= W; chosen for the shared-memory bank mapping)
  int* work;     // dynamic scheduling: global work-item counter (zeroed by the caller); null = static round-robin
  double* acc;
};

// Barrier-free dynamic work distribution inside a CTA.  All lane groups of a CTA must walk the SAME sequence of work
// items (they own different rows of the same tile), but they drift apart freely.  The sequence is therefore kept in
// a small shared-memory ring: entry `it` is produced by whichever group gets there first (it claims the entry with an
// atomicCAS on `head`, takes the next item from the global counter and publishes {it + 1, item} with one 64-bit
// store); the others read it.  `cnt[slot]` counts the reads of a ring slot, so that the producer of entry it + NSQ
// only overwrites entry `it` once every group of the CTA has consumed it.
constexpr int NSQ = 32;
struct DynSched {
  unsigned long long seq[NSQ];
  int cnt[NSQ];
  int head;
  int pad;
};

__device__ __forceinline__ double lds_f64(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }

// A lane group (S lanes) owns R consecutive rows of the shared-memory tile: group G of a window works on matrix
// strip (r0 / R) + G and adds pixel (q, col) to tile row G * R + q.  Tile row t holds window row di = t - (r0 mod R);
// r0 mod R is constant inside an extended slot, so groups never write another group's rows, the read-modify-write
// `tile[t][col - c0] += val` needs no atomics, and -- because a group also flushes (`red.global.add.f64`) and
// clears its own rows when the extended slot changes -- the kernel has no CTA-wide barrier at all: warps drift
// freely through the CTA's (static, round-robin) chunk list.  Tile rows whose di falls outside [0, W) (the part of
// the first / last strip above / below the window) collect pixels that are simply never flushed.
// WU windows advance in lockstep per group (two register sets, ping-pong), so WU independent 16-byte loads are in
// flight per lane while the previous S pixels of every window are added to the tile.  A group's load covers
// S * 16 contiguous, aligned bytes: with S >= 8 every warp-wide load touches 4 full 128-byte lines.
template <int R, int S, int WU, int PF, int MINB, bool QI>
__global__ void __launch_bounds__(S == 32 ? NT_MAX : NT_MAX - 32, MINB) k_pileup_main(const MainParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LR = (R == 1) ? 0 : (R == 2) ? 1 : (R == 4) ? 2 : 3;
  const int W = p.W;
  const int TW = p.TW;
  const AccLayout L(W);
  const int lane = threadIdx.x & 31;
  const int sub = lane / S;
  const int ls = lane % S;
  const int g = (threadIdx.x >> 5) * (32 / S) + sub;  // group id inside the band
  const unsigned gmask = (S == 32) ? 0xffffffffu : (((1u << S) - 1u) << (sub * S));
  const int ns = p.ns;
  DynSched* ds = reinterpret_cast<DynSched*>(smem_raw + (size_t)p.Gb * R * TW * 8);
  if (p.work != nullptr) {  // the only CTA-wide barrier of the kernel: the scheduler ring starts empty
    for (int i = threadIdx.x; i < (int)(sizeof(DynSched) / 4); i += blockDim.x) reinterpret_cast<int*>(ds)[i] = 0;
    __syncthreads();
  }
  if (g >= p.Gb) return;  // no barriers below: idle groups may leave

  unsigned trow = (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)(g * R * TW) * 8u;  // my R tile rows
  asm volatile("mov.u32 %0, %0;" : "+r"(trow));  // keep the address in a register (no rematerialisation)
  for (int i = ls; i < R * TW; i += S) sts_f64(trow + i * 8, 0.0);
  __syncwarp(gmask);
  int cur_slot = -1, cur_band = 0;  // extended slot
  const int total_chunks = __ldg(&p.chunks.chunk_start[p.chunks.n_slots]);
  const int total_items = total_chunks * p.n_bands;

  // cell (q, dj) of my rows: row-major with stride TW, or (QI) the R rows interleaved column by column
  auto cell = [&](int q, int dj) -> unsigned {
    return QI ? trow + (unsigned)(dj * R + q) * 8u : trow + (unsigned)(q * TW + dj) * 8u;
  };

  auto flush_rows = [&]() {
    // this group's tile rows -> global accumulator (red.global.add.f64), then clear them
    __syncwarp(gmask);
    const int m = cur_slot & (R - 1);
    const int t0 = (cur_band * p.Gb + g) * R;
    double* abase = p.acc + (int64_t)(cur_slot >> LR) * L.stride;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int di = t0 + q - m;
      if ((unsigned)di < (unsigned)W) {
        double* a = abase + (int64_t)di * W;
        for (int dj = ls; dj < W; dj += S) {
          const unsigned ad = cell(q, dj);
          const double v = lds_f64(ad);
          if (v != 0.0) {
            atomicAdd(a + dj, v);
            sts_f64(ad, 0.0);
          }
        }
      } else {
        for (int dj = ls; dj < W; dj += S) sts_f64(cell(q, dj), 0.0);
      }
    }
    __syncwarp(gmask);
  };

  // next work item of this CTA's sequence (static: blockIdx.x, + gridDim.x, ...; dynamic: see DynSched)
  int it = 0;
  auto next_item = [&]() -> int {
    if (p.work == nullptr) return blockIdx.x + (it++) * gridDim.x;
    int item = 0;
    if (ls == 0) {
      const int slot = it & (NSQ - 1);
      volatile unsigned long long* sq = &ds->seq[slot];
      volatile int* head = &ds->head;
      for (;;) {
        const unsigned long long v = *sq;
        if ((int)(v >> 32) == it + 1) {
          item = (int)(unsigned)v;
          break;
        }
        if (*head == it && atomicCAS(&ds->head, it, it + 1) == it) {
          // I produce entry `it`: wait until all groups have read the previous occupant of the ring slot
          const int need = p.Gb * (it / NSQ);
          while (*(volatile int*)&ds->cnt[slot] < need) {
          }
          item = atomicAdd(p.work, 1);
          if (item > total_items) item = total_items;
          *sq = ((unsigned long long)(unsigned)(it + 1) << 32) | (unsigned)item;
          break;
        }
      }
      atomicAdd(&ds->cnt[slot], 1);
    }
    ++it;
    return __shfl_sync(gmask, item, sub * S);
  };

  for (int item = next_item(); item < total_items; item = next_item()) {
    const int band = item / total_chunks;
    int slot, w_lo, w_hi;
    locate_chunk(p.chunks, item - band * total_chunks, slot, w_lo, w_hi);
    if (slot != cur_slot || band != cur_band) {
      if (cur_slot >= 0) flush_rows();
      cur_slot = slot;
      cur_band = band;
    }
    const int G = band * p.Gb + g;  // my strip of every window of this chunk
    if (G >= p.n_groups || G * R - (slot & (R - 1)) >= W) continue;  // strip entirely below the window

    for (int w = w_lo; w < w_hi; w += WU) {
      // Two register sets (A, B) of raw pixel records {col, q, val} per window in flight.  Every half-step loads
      // the next record of every run with a predicated load that lands directly in the other register set; a lane
      // whose current pixel already lies right of the window gets col = INT_MAX instead, which ends its run.
      int idx[WU], c0s[WU];
      int4 A[WU], B[WU];
      // stage A: window records -> strip pointers (WU independent chains)
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        idx[u] = ls;  // a sentinel
        c0s[u] = 0;
        if (w + u < w_hi) {
          const int2 rc = __ldg(&p.win[w + u]);
          idx[u] = __ldg(&p.bucket[(rc.y >> p.lb) * ns + (rc.x >> LR) + G]) + ls;
          c0s[u] = rc.y;
        }
      }
      // stage B: first pixel of every window's run
#pragma unroll
      for (int u = 0; u < WU; ++u) A[u] = __ldg(reinterpret_cast<const int4*>(p.pix + idx[u]));
      // stage C: the WU runs advance in lockstep, S pixels per run per half-step.  Strips end with a full group of
      // sentinel pixels, so "dj >= W" is the only termination test.
#define PUP_HALF_STEP(C, N)                                                                      \
  {                                                                                              \
    int dj[WU];                                                                                  \
    _Pragma("unroll") for (int u = 0; u < WU; ++u) dj[u] = C[u].x - c0s[u];                      \
    int mn = dj[0];                                                                              \
    _Pragma("unroll") for (int u = 1; u < WU; ++u) mn = min(mn, dj[u]);                          \
    if (mn >= W) break;                                                                          \
    _Pragma("unroll") for (int u = 0; u < WU; ++u) {                                             \
      idx[u] += S;                                                                               \
      if (dj[u] < W) {                                                                           \
        const Pix* src = p.pix + idx[u];                                                         \
        N[u] = __ldg(reinterpret_cast<const int4*>(src));                                        \
        if (PF == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 2 * S));                \
      } else {                                                                                   \
        N[u].x = 0x7fffffff;                                                                     \
      }                                                                                          \
    }                                                                                            \
    _Pragma("unroll") for (int u = 0; u < WU; ++u) {                                             \
      if ((unsigned)dj[u] < (unsigned)W) {                                                       \
        const unsigned a = cell(R > 1 ? C[u].y : 0, dj[u]);                                      \
        sts_f64(a, lds_f64(a) + __hiloint2double(C[u].w, C[u].z));                               \
      }                                                                                          \
    }                                                                                            \
  }
      // Lanes leave the loop individually once their WU runs are exhausted and wait at the __syncwarp below; the
      // lanes still inside execute the same predicated instruction stream, so the read-modify-writes of one tile
      // row are ordered by the program order of a converged SIMT group (racecheck reports them as intra-warp
      // hazards "without barrier"; a group-voted exit was measured 6 % slower and changes nothing about ordering).
      // Two lanes of a group never hold the same tile cell in the same half-step: the S pixels a group loads per
      // window and half-step are S distinct (col, q) entries of one strip, and the WU windows of a half-step are
      // added one after the other (pinned by tests/test_gpu_parity.py::test_adversarial_duplicate_windows and
      // the SASS excerpt in profiles/).
      for (;;) {
        PUP_HALF_STEP(A, B)
        PUP_HALF_STEP(B, A)
      }
#undef PUP_HALF_STEP
      __syncwarp(gmask);
    }
  }
  if (cur_slot >= 0) flush_rows();
}


// ------------------------------------------------------------------------------------------ dense-band pile-up
// Windows close to the diagonal are dense (in the synthetic 3 Gbp genome a window 10 Mb off the diagonal still holds
// a pixel in 40 % of its cells): for them a sparse scatter -- one 16-byte record load plus one fp64 shared-memory
// read-modify-write per stored pixel, ~0.55 L1 wavefronts per pixel -- costs more than reading every cell of the
// window from a dense copy of the matrix and adding it to a REGISTER: thread t owns the same cells (t, t + T, ...) of
// the W x W tile for every window, so there is no scatter, no shared memory and no atomics until the slot changes.
// The dense copy is a diagonal band, band[row][col - row] for col - row < bw (built by k_band_* at region creation,
// bw chosen on the device where the pixel density falls below PUP_BAND_DENSITY_PCT); a window row is W consecutive
// doubles of one band row, so a warp's load is one contiguous 256-byte run (~3 L1 wavefronts per 32 cells).
// Break-even against the sparse path: ~18 % occupancy in L1 wavefronts.  Windows of the class (k_window_keys) come
// sorted by (slot, r0): the CTAs of the grid walk neighbouring band rows, which stay in L2.
constexpr int DENSE_BAND_CELLS = 1024 * 7;  // tiles with more cells are cut into row bands of at most this many

struct DenseParams {
  int W, stride, lr, n_slots;
  int rows_per_band, n_bands;  // the tile's rows are piled up in n_bands passes over the window list (W = 203: 6)
  const double* band;
  const int2* win;    // sorted (r0, c0)
  ChunkTable chunks;  // over all extended slots of both classes
  int first_eslot;    // first extended slot of the dense class
  int* work;
  double* acc;
};

#ifndef PUP_DENSE_LD
#define PUP_DENSE_LD __ldg  // tuning: -DPUP_DENSE_LD=__ldcg / __ldcs / __ldlu
#endif
template <int DENSE_T, int DENSE_CPT>
__global__ void __launch_bounds__(DENSE_T, 1024 / DENSE_T) k_pileup_dense(const DenseParams p) {
  __shared__ int s_item;
  const int W = p.W;
  const AccLayout L(W);
  const int t = threadIdx.x;
  // band element of my cell k relative to the window's first element; cells beyond the tile read element 0 of the
  // window (a valid address) into an accumulator that is never flushed: the inner loop carries no predicates
  unsigned off[DENSE_CPT];
  int cellid[DENSE_CPT];
  double a[DENSE_CPT];
  const double* __restrict__ bandp = p.band;
#pragma unroll
  for (int k = 0; k < DENSE_CPT; ++k) {
    off[k] = 0u;
    cellid[k] = -1;
    a[k] = 0.0;
  }
  const int c_first = __ldg(&p.chunks.chunk_start[p.first_eslot]);
  const int n_chunks = __ldg(&p.chunks.chunk_start[p.chunks.n_slots]) - c_first;
  const int n_items = n_chunks * p.n_bands;
  int cur = -1, cur_band = -1;
  auto flush = [&]() {
    if (cur < 0) return;
    double* dst = p.acc + (int64_t)cur * L.stride;
#pragma unroll
    for (int k = 0; k < DENSE_CPT; ++k) {
      if (cellid[k] >= 0 && a[k] != 0.0) atomicAdd(dst + cellid[k], a[k]);
      a[k] = 0.0;
    }
  };
  for (;;) {
    __syncthreads();
    if (t == 0) s_item = atomicAdd(p.work, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int band = item / n_chunks;
    int es, lo, hi;
    locate_chunk(p.chunks, c_first + (item - band * n_chunks), es, lo, hi);
    const int slot = (es >> p.lr) - p.n_slots;
    if (slot != cur || band != cur_band) {
      flush();
      cur = slot;
      if (band != cur_band) {  // my cells of this band: band-local index t, t + T, ... -> (row, column) of the tile
        cur_band = band;
        const int i0 = band * p.rows_per_band, rows = min(p.rows_per_band, W - i0);
#pragma unroll
        for (int k = 0; k < DENSE_CPT; ++k) {
          const int idx = t + k * DENSE_T;
          const int il = idx / W, j = idx - il * W, i = i0 + il;
          const bool ok = il < rows;
          off[k] = ok ? (unsigned)(i * (p.stride - 1) + j) : 0u;  // (r0 + i) * stride + (c0 + j) - (r0 + i) - window base
          cellid[k] = ok ? i * W + j : -1;
        }
      }
    }
    // two windows in flight; 32-bit element indices (the band holds < 2^31 doubles): one IMAD.WIDE per address
    int w = lo;
    for (; w + 1 < hi; w += 2) {
      const int2 rc0 = __ldg(&p.win[w]);
      const int2 rc1 = __ldg(&p.win[w + 1]);
      const unsigned e0 = (unsigned)(rc0.x * p.stride + (rc0.y - rc0.x));
      const unsigned e1 = (unsigned)(rc1.x * p.stride + (rc1.y - rc1.x));
      double v0[DENSE_CPT], v1[DENSE_CPT];
#pragma unroll
      for (int k = 0; k < DENSE_CPT; ++k) v0[k] = PUP_DENSE_LD(bandp + (e0 + off[k]));
#pragma unroll
      for (int k = 0; k < DENSE_CPT; ++k) v1[k] = PUP_DENSE_LD(bandp + (e1 + off[k]));
#pragma unroll
      for (int k = 0; k < DENSE_CPT; ++k) {
        a[k] += v0[k];
        a[k] += v1[k];
      }
    }
    if (w < hi) {
      const int2 rc0 = __ldg(&p.win[w]);
      const unsigned e0 = (unsigned)(rc0.x * p.stride + (rc0.y - rc0.x));
#pragma unroll
      for (int k = 0; k < DENSE_CPT; ++k) a[k] += PUP_DENSE_LD(bandp + (e0 + off[k]));
    }
  }
  flush();
}

// ---- building the band (region creation)
constexpr int BAND_HB = 4096;  // histogram blocks of 2^hs diagonals each

// pixels per block of diagonals (one warp per strip; sentinels have col = INT_MAX)
__global__ void __launch_bounds__(256) k_band_hist(const Pix* __restrict__ pix, const int32_t* __restrict__ prow, int ns,
                                                    int lr, int hs, int* __restrict__ hist) {
  __shared__ int sh[BAND_HB];
  for (int i = threadIdx.x; i < BAND_HB; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int s = warp; s < ns; s += nwarps) {
    const int lo = prow[s], hi = prow[s + 1];
    for (int k = lo + lane; k < hi; k += 32) {
      const int2 cq = __ldg(reinterpret_cast<const int2*>(pix + k));
      if (cq.x == 0x7fffffff) continue;
      const int d = cq.x - ((s << lr) + cq.y);
      if (d >= 0 && (d >> hs) < BAND_HB) atomicAdd(&sh[d >> hs], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BAND_HB; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// bw = end of the last block of diagonals (from the diagonal outwards) whose occupancy is still >= pct percent
__global__ void k_band_choose(const int* __restrict__ hist, int nb, int hs, int bw_max, int pct, int ignore_diags,
                              int32_t* __restrict__ bw_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int bw = 0;
  for (int b = 0; (b << hs) < nb && b < BAND_HB; ++b) {
    long long d0 = (long long)b << hs;
    const long long d1 = min((long long)nb, d0 + (1ll << hs));
    d0 = max(d0, (long long)ignore_diags);  // the masked diagonals hold nothing by construction: not counted
    if (d0 >= d1) {
      bw = (int)d1;
      continue;
    }
    const long long cells = (d1 - d0) * nb - (d0 + d1 - 1) * (d1 - d0) / 2;  // positions (row, row + d), d in [d0, d1)
    if ((long long)hist[b] * 100 < cells * pct) break;
    bw = (int)d1;
  }
  bw = min(bw, bw_max);
  *bw_out = bw >= 32 ? bw : 0;
}

__global__ void k_band_zero(double2* __restrict__ band, int nb, int stride, const int32_t* __restrict__ bw_ptr) {
  const int half = stride >> 1;  // stride is a multiple of 16
  const int hw = (__ldg(bw_ptr) + 1) >> 1;
  const int64_t total = (int64_t)nb * hw;
  if (hw == 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / hw;
    band[row * half + (i - row * hw)] = make_double2(0.0, 0.0);
  }
}

__global__ void __launch_bounds__(256) k_band_fill(const Pix* __restrict__ pix, const int32_t* __restrict__ prow, int ns,
                                                    int lr, double* __restrict__ band, int stride,
                                                    const int32_t* __restrict__ bw_ptr) {
  const int bw = __ldg(bw_ptr);
  if (bw == 0) return;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int s = warp; s < ns; s += nwarps) {
    const int lo = prow[s], hi = prow[s + 1];
    for (int k = lo + lane; k < hi; k += 32) {
      const int4 raw = __ldg(reinterpret_cast<const int4*>(pix + k));
      if (raw.x == 0x7fffffff) continue;
      const int row = (s << lr) + raw.y, d = raw.x - row;
      if (d >= 0 && d < bw) band[(int64_t)row * stride + d] = __hiloint2double(raw.w, raw.z);
    }
  }
}

// ------------------------------------------------------------------------------------------ dense-num kernel
// `num` of slow windows: every pixel of the window is tested (row / column weight, signed diagonal, expected).
// One CTA per chunk; thread t owns tile cells t, t + blockDim, ... so the int32 tile needs no atomics.
struct SlowParams {
  WinCtx ctx;
  const uint8_t* bad;   // null for raw
  const uint8_t* ebad;
  const uint64_t* keys;
  ChunkTable chunks;
  double* acc;
  const int* n_slow;
  int* counter;
};

__global__ void __launch_bounds__(512) k_num_slow(const SlowParams p) {
  if (*p.n_slow == 0) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* numT = reinterpret_cast<int*>(smem_raw);
  __shared__ int s_slot, s_lo, s_hi;
  const int W = p.ctx.W;
  const int w2 = W * W;
  const AccLayout L(W);
  const bool nodiag = p.ctx.flags & PUP_F_NODIAG;
  const bool ooe = p.ctx.flags & PUP_F_OOE;
  const int igd = p.ctx.ignore_diags;
  for (int i = threadIdx.x; i < w2; i += blockDim.x) numT[i] = 0;
  int cur_slot = -1;
  const int total_chunks = __ldg(&p.chunks.chunk_start[p.chunks.n_slots]);
  auto flush = [&]() {
    double* a = p.acc + (int64_t)cur_slot * L.stride + L.off_num;  // cur_slot: accumulator slot
    for (int i = threadIdx.x; i < w2; i += blockDim.x) {
      int c = numT[i];
      if (c != 0) {
        atomicAdd(a + i, (double)c);
        numT[i] = 0;
      }
    }
  };
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int item = atomicAdd(p.counter, 1);
      if (item >= total_chunks) {
        s_slot = -1;
      } else {
        int slot, lo, hi;
        locate_chunk(p.chunks, item, slot, lo, hi);
        s_slot = slot >> p.ctx.lr;  // extended slot -> accumulator slot
        if (s_slot >= p.ctx.n_slots) s_slot -= p.ctx.n_slots;
        s_lo = lo;
        s_hi = hi;
      }
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot < 0) break;
    if (slot != cur_slot) {
      if (cur_slot >= 0) {
        flush();
        __syncthreads();
      }
      cur_slot = slot;
    }
    for (int w = s_lo; w < s_hi; ++w) {
      int kslot, r0, c0;
      decode_key(__ldg(&p.keys[w]), p.ctx, kslot, r0, c0);
      if (!window_is_slow(p.ctx, r0, c0)) continue;
      for (int cell = threadIdx.x; cell < w2; cell += blockDim.x) {
        const int di = cell / W, dj = cell - di * W;
        const int r = r0 + di, c = c0 + dj;
        const int d = c - r;
        bool ok = true;
        if (p.bad != nullptr) ok = !__ldg(&p.bad[r]) && !__ldg(&p.bad[c]);
        if (!nodiag) ok = ok && (d >= igd);
        if (ooe) ok = ok && !__ldg(&p.ebad[d < 0 ? -d : d]);
        if (ok) numT[cell] += 1;
      }
    }
  }
  __syncthreads();
  if (cur_slot >= 0) flush();
}

// ------------------------------------------------------------------------------------------ stripes
// Per-window centre row / centre column (store_stripes, coolpup.py:1164-1169): a per-ROI output, not a reduction.
// One warp per window; every lane looks its pixel up by binary search in the (unnormalised-NaN-aware) row.
struct StripeParams {
  const Pix* pix;
  const int32_t* prow;      // strip starts
  const uint8_t* bad;       // null for raw
  const double* expected;   // null unless the region was prepared with PUP_F_OOE
  int nb, W, ignore_diags, lr;
  unsigned flags;
};

__device__ __forceinline__ double snippet_pixel(const StripeParams& p, int r, int c) {
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (p.bad != nullptr && (p.bad[r] || p.bad[c])) return nan;
  const int d = c - r;
  if (!(p.flags & PUP_F_NODIAG) && d < p.ignore_diags) return nan;
  const int s = r >> p.lr, q = r & ((1 << p.lr) - 1);
  int lo = p.prow[s], hi = p.prow[s + 1];  // (col, q)-sorted, closed by sentinels
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    const int2 k = __ldg(reinterpret_cast<const int2*>(p.pix + mid));
    if (k.x < c || (k.x == c && k.y < q))
      lo = mid + 1;
    else
      hi = mid;
  }
  bool stored = false;
  if (lo < p.prow[s + 1]) {
    const int2 k = __ldg(reinterpret_cast<const int2*>(p.pix + lo));
    stored = k.x == c && k.y == q;
  }
  const double v = stored ? p.pix[lo].val : 0.0;
  if ((p.flags & PUP_F_OOE) && p.expected != nullptr) {
    const double e = p.expected[d < 0 ? -d : d];
    if (isnan(e)) return nan;
    if (e == 0.0) return (stored && isinf(v)) ? v : nan;  // x/0 = inf, 0/0 = NaN
  }
  return v;
}

__global__ void k_stripes(const StripeParams p, const int32_t* __restrict__ r0, const int32_t* __restrict__ c0,
                          int64_t n, double* __restrict__ hor, double* __restrict__ ver) {
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const int W = p.W, cn = W / 2;
  const int r = r0[w], c = c0[w];
  const bool ok = r >= 0 && c >= 0 && r + W <= p.nb && c + W <= p.nb;
  for (int k = lane; k < W; k += 32) {
    hor[w * W + k] = ok ? snippet_pixel(p, r + cn, c + k) : nan;           // data[cntr, :]
    ver[w * W + k] = ok ? snippet_pixel(p, r + (W - 1 - k), c + cn) : nan;  // data[:, cntr][::-1]
  }
}

// ------------------------------------------------------------------------------------------ rescaled pile-ups
// PileUpper._rescale_snip (coolpup.py:1193-1234): a window of its feature's own size h x w is zoomed to rs x rs with
// cooltools' zoom_array -- scipy.ndimage.zoom(order=1) to the next multiple of rs, then block means -- once for the
// snippet with NaN -> 0 and once for its NaN mask; output cells that any NaN touches become NaN.  The arithmetic of
// scipy's NI_ZoomShift is mirrored exactly (validated bit for bit against scipy on the CPU, oracle/zoom_ref.py, tests/test_zoom_ref.py):
//   out size n_tmp = rs * mult, mult = ceil(n / rs) when n > rs else 1;  coordinate cc = k * ((n - 1) / (n_tmp - 1));
//   cc > n - 1 (rounding) -> the sample is the constant 0;  i0 = floor(cc), i1 = i0 + 1 mirrored at the edge
//   (2n - 2 - i1), weights w0 = 1 - (cc - i0), w1 = 1 - w0;  sample = ((D[i0,j0]*wr0)*wc0 + (D[i0,j1]*wr0)*wc1) + ...
struct ZoomPlan {
  int i0, i1;
  double w0, w1;  // both 0: the sample is the constant 0
};

__device__ __forceinline__ ZoomPlan zoom_plan_entry(int k, int n, int n_tmp) {
  ZoomPlan z;
  z.i0 = z.i1 = 0;
  z.w0 = z.w1 = 0.0;
  const double zf = n_tmp > 1 ? __ddiv_rn((double)(n - 1), (double)(n_tmp - 1)) : 1.0;
  const double cc = __dmul_rn((double)k, zf);
  if (cc < 0.0 || cc > (double)(n - 1)) return z;
  const double fl = floor(cc);
  const double x = __dsub_rn(cc, fl);
  z.w0 = __dsub_rn(1.0, x);
  z.w1 = __dsub_rn(1.0, z.w0);
  if (n > 1) {
    z.i0 = (int)fl;
    z.i1 = z.i0 + 1;
    if (z.i1 >= n) z.i1 = 2 * n - 2 - z.i1;
  }
  return z;
}

struct RescaleParams {
  StripeParams sp;         // snippet semantics of the region (same as pup_stripes)
  const double* expected;  // raw expected vector (mode-1 windows: the bare expected block as a snippet)
  const double* coverage;  // or null
  const int32_t *r0, *c0, *h, *w, *slot, *mode;
  int64_t n_win;
  int rs, n_slots, max_tmp;
  unsigned flags;          // PUP_F_LOCAL, PUP_F_COVERAGE
  double* scratch_d;       // [grid][max_cells]
  uint8_t* scratch_m;      // [grid][max_cells]
  int64_t max_cells;
  double* acc;
  int* work;               // [0] next window, [1] in-bounds windows
};

__global__ void __launch_bounds__(1024) k_rescale(const RescaleParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rs = p.rs, rs2 = rs * rs;
  double* tsum = reinterpret_cast<double*>(smem_raw);
  int* tnum = reinterpret_cast<int*>(tsum + rs2);
  ZoomPlan* prow = reinterpret_cast<ZoomPlan*>(smem_raw + (((size_t)rs2 * 12 + 15) / 16) * 16);
  ZoomPlan* pcol = prow + p.max_tmp;
  __shared__ int s_win;
  __shared__ int s_n;
  const AccLayout L(rs);
  const int tid = threadIdx.x, nt = blockDim.x;
  double* D = p.scratch_d + (size_t)blockIdx.x * p.max_cells;
  uint8_t* M = p.scratch_m + (size_t)blockIdx.x * p.max_cells;
  for (int i = tid; i < rs2; i += nt) {
    tsum[i] = 0.0;
    tnum[i] = 0;
  }
  if (tid == 0) s_n = 0;
  int cur_slot = -1;
  auto flush = [&]() {
    __syncthreads();
    if (cur_slot >= 0) {
      double* a = p.acc + (int64_t)cur_slot * L.stride;
      for (int i = tid; i < rs2; i += nt) {
        if (tsum[i] != 0.0) atomicAdd(a + i, tsum[i]);
        if (tnum[i] != 0) atomicAdd(a + L.off_num + i, (double)tnum[i]);
        tsum[i] = 0.0;
        tnum[i] = 0;
      }
      if (tid == 0 && s_n) {
        atomicAdd(a + L.off_n, (double)s_n);
        s_n = 0;
      }
    }
    __syncthreads();
  };
  for (;;) {
    __syncthreads();
    if (tid == 0) s_win = atomicAdd(p.work, 1);
    __syncthreads();
    const int wi = s_win;
    if (wi >= p.n_win) break;
    const int r = p.r0[wi], c = p.c0[wi], hh = p.h[wi], ww = p.w[wi], sl = p.slot[wi];
    const int md = p.mode ? p.mode[wi] : 0;
    if (r < 0 || c < 0 || hh < 0 || ww < 0 || r + hh > p.sp.nb || c + ww > p.sp.nb || sl < 0 || sl >= p.n_slots) continue;
    if (sl != cur_slot) {
      flush();
      cur_slot = sl;
    }
    if (tid == 0) {
      s_n += 1;
      atomicAdd(p.work + 1, 1);
    }
    const int cells = hh * ww;
    // 1. the dense snippet exactly as _stream_snips builds it: first what an UNSTORED pixel is -- 0, or NaN under a
    //    masked bin, a masked diagonal, a NaN expected or 0 / 0 -- then the stored pixels of the window's strips,
    //    streamed once (a per-cell lookup would binary-search the strip h * w times)
    bool some = false;
    if (md == 1) {
      for (int idx = tid; idx < cells; idx += nt) {
        const int di = idx / ww, dj = idx - di * ww;
        const int d = (c + dj) - (r + di);
        const double v = p.expected[d < 0 ? -d : d];
        D[idx] = v;
        some |= !isnan(v);
      }
    } else {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      const bool ooe = (p.sp.flags & PUP_F_OOE) && p.sp.expected != nullptr;
      const bool nodiag = p.sp.flags & PUP_F_NODIAG;
      for (int idx = tid; idx < cells; idx += nt) {
        const int di = idx / ww, dj = idx - di * ww;
        const int row = r + di, col = c + dj, d = col - row;
        double v = 0.0;
        if (p.sp.bad != nullptr && (p.sp.bad[row] || p.sp.bad[col])) {
          v = nan;
        } else if (!nodiag && d < p.sp.ignore_diags) {
          v = nan;
        } else if (ooe) {
          const double e = p.sp.expected[d < 0 ? -d : d];
          if (isnan(e) || e == 0.0) v = nan;
        }
        D[idx] = v;
        some |= !isnan(v);
      }
      __syncthreads();
      if (cells > 0) {
        const int lane = tid & 31, nwarps = nt >> 5;
        const int s_hi = (r + hh - 1) >> p.sp.lr;
        for (int s = (r >> p.sp.lr) + (tid >> 5); s <= s_hi; s += nwarps) {
          int a = p.sp.prow[s], b = p.sp.prow[s + 1];
          const int end = b;
          while (a < b) {  // first record of the strip with col >= c (the same search in every lane)
            const int mid = (a + b) >> 1;
            if (__ldg(&p.sp.pix[mid].col) < c)
              a = mid + 1;
            else
              b = mid;
          }
          for (int k = a + lane; k < end; k += 32) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(p.sp.pix + k));
            const int col = raw.x;
            if (col >= c + ww) break;  // the sentinels that close a strip end the run too
            const int row = (s << p.sp.lr) + raw.y, di = row - r;
            if (di < 0 || di >= hh) continue;
            if (p.sp.bad != nullptr && (p.sp.bad[row] || p.sp.bad[col])) continue;
            const int d = col - row;
            if (!nodiag && d < p.sp.ignore_diags) continue;
            const double v = __hiloint2double(raw.w, raw.z);
            if (ooe) {
              const double e = p.sp.expected[d < 0 ? -d : d];
              if (isnan(e)) continue;
              if (e == 0.0 && !isinf(v)) continue;  // x / 0 = inf, 0 / 0 = NaN
            }
            D[di * ww + (col - c)] = v;
            some = true;
          }
        }
      }
    }
    const int any = __syncthreads_or(some ? 1 : 0);
    const int mr = hh > rs ? (hh + rs - 1) / rs : 1, mc = ww > rs ? (ww + rs - 1) / rs : 1;
    if (cells == 0 || !any) {  // size 0 or all NaN: the snippet is a block of zeros (coolpup.py:1212-1213)
      for (int i = tid; i < rs2; i += nt) tnum[i] += 1;
      if (!(p.flags & PUP_F_COVERAGE) || cells == 0) continue;
      // the coverage vectors are zoomed all the same (1229-1233)
      for (int k = tid; k < rs * mr; k += nt) prow[k] = zoom_plan_entry(k, hh, rs * mr);
      for (int k = tid; k < rs * mc; k += nt) pcol[k] = zoom_plan_entry(k, ww, rs * mc);
      __syncthreads();
    } else {
    // 2. local pile-ups are symmetrised before the zoom (1215-1220): nanmean of the snippet and its transpose
    if ((p.flags & PUP_F_LOCAL) && hh == ww) {
      for (int idx = tid; idx < cells; idx += nt) {
        const int di = idx / ww, dj = idx - di * ww;
        if (di < dj) {
          const double a = D[idx], b = D[dj * ww + di];
          double m;
          if (isnan(a))
            m = b;
          else if (isnan(b))
            m = a;
          else
            m = __ddiv_rn(__dadd_rn(a, b), 2.0);
          D[idx] = m;
          D[dj * ww + di] = m;
        }
      }
      __syncthreads();
    }
    // 3. NaN mask, nan_to_num (NaN -> 0, +-inf -> +-DBL_MAX), zoom plans of both axes
    for (int idx = tid; idx < cells; idx += nt) {
      const double v = D[idx];
      const bool isn = isnan(v);
      M[idx] = isn ? 1 : 0;
      if (isn)
        D[idx] = 0.0;
      else if (isinf(v))
        D[idx] = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    }
    for (int k = tid; k < rs * mr; k += nt) prow[k] = zoom_plan_entry(k, hh, rs * mr);
    for (int k = tid; k < rs * mc; k += nt) pcol[k] = zoom_plan_entry(k, ww, rs * mc);
    __syncthreads();
    // 4. every output cell: block mean (rows first, then columns, like np.mean over the split axes) of the
    //    interpolated samples, for the data and for the mask
    for (int cell = tid; cell < rs2; cell += nt) {
      const int i = cell / rs, j = cell - i * rs;
      double acc_d = 0.0, acc_m = 0.0;
      for (int bj = 0; bj < mc; ++bj) {
        const ZoomPlan zc = pcol[j * mc + bj];
        double col_d = 0.0, col_m = 0.0;
        for (int bi = 0; bi < mr; ++bi) {
          const ZoomPlan zr = prow[i * mr + bi];
          double t = 0.0, tm = 0.0;
          if ((zr.w0 != 0.0 || zr.w1 != 0.0) && (zc.w0 != 0.0 || zc.w1 != 0.0)) {
            const int a0 = zr.i0 * ww, a1 = zr.i1 * ww;
            t = __dmul_rn(__dmul_rn(D[a0 + zc.i0], zr.w0), zc.w0);
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(D[a0 + zc.i1], zr.w0), zc.w1));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(D[a1 + zc.i0], zr.w1), zc.w0));
            t = __dadd_rn(t, __dmul_rn(__dmul_rn(D[a1 + zc.i1], zr.w1), zc.w1));
            tm = __dmul_rn(__dmul_rn((double)M[a0 + zc.i0], zr.w0), zc.w0);
            tm = __dadd_rn(tm, __dmul_rn(__dmul_rn((double)M[a0 + zc.i1], zr.w0), zc.w1));
            tm = __dadd_rn(tm, __dmul_rn(__dmul_rn((double)M[a1 + zc.i0], zr.w1), zc.w0));
            tm = __dadd_rn(tm, __dmul_rn(__dmul_rn((double)M[a1 + zc.i1], zr.w1), zc.w1));
          }
          col_d = __dadd_rn(col_d, t);
          col_m = __dadd_rn(col_m, tm);
        }
        if (mr > 1) {
          col_d = __ddiv_rn(col_d, (double)mr);
          col_m = __ddiv_rn(col_m, (double)mr);
        }
        acc_d = __dadd_rn(acc_d, col_d);
        acc_m = __dadd_rn(acc_m, col_m);
      }
      if (mc > 1) {
        acc_d = __ddiv_rn(acc_d, (double)mc);
        acc_m = __ddiv_rn(acc_m, (double)mc);
      }
      if (!(acc_m > 0.0)) {  // np.ceil(nanzoom).astype(bool): any NaN weight makes the cell NaN
        if (isfinite(acc_d)) {
          tsum[cell] += acc_d;
          tnum[cell] += 1;
        } else if (!isnan(acc_d)) {
          tsum[cell] += acc_d;  // an infinite cell poisons the sum like in the reference (not counted in num)
        }
      }
    }
    }
    // 5. coverage vectors are zoomed the same way in one dimension (1229-1233)
    if ((p.flags & PUP_F_COVERAGE) && p.coverage != nullptr) {
      double* a = p.acc + (int64_t)sl * L.stride;
      for (int o = tid; o < 2 * rs; o += nt) {
        const bool row = o < rs;
        const int i = row ? o : o - rs;
        const int n = row ? hh : ww, m = row ? mr : mc, base = row ? r : c;
        const ZoomPlan* pl = row ? prow : pcol;
        double v = 0.0;
        for (int b = 0; b < m; ++b) {
          const ZoomPlan z = pl[i * m + b];
          double t = 0.0;
          if (z.w0 != 0.0 || z.w1 != 0.0)
            t = __dadd_rn(__dmul_rn(p.coverage[base + z.i0], z.w0), __dmul_rn(p.coverage[base + z.i1], z.w1));
          v = __dadd_rn(v, t);
        }
        if (m > 1) v = __ddiv_rn(v, (double)m);
        (void)n;
        if (!isnan(v) && v != 0.0) atomicAdd(a + (row ? L.off_covs : L.off_cove) + i, v);
      }
    }
  }
  flush();
}

// ------------------------------------------------------------------------------------------ byte counter
// Exact algorithmic pixel count of a window list (measurement helper, not on the timed path).
__global__ void k_count_nnz(const Pix* __restrict__ pix, const int32_t* __restrict__ prow,
                            const int32_t* __restrict__ r0, const int32_t* __restrict__ c0, int64_t n, int nb, int W,
                            int lr, unsigned long long* out_nnz, unsigned long long* out_valid) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int R = 1 << lr;
  unsigned long long local = 0, valid = 0;
  for (int64_t i = warp; i < n; i += nwarps) {
    int r = r0[i], c = c0[i];
    if (r < 0 || c < 0 || r + W > nb || c + W > nb) continue;
    if (lane == 0) valid += 1;
    const int ng = ((r & (R - 1)) + W + R - 1) >> lr;  // strips the window touches
    for (int G = lane; G < ng; G += 32) {
      const int s = (r >> lr) + G;
      const int lo = prow[s], hi = prow[s + 1];
      auto lower = [&](int target) {
        int a = lo, b = hi;
        while (a < b) {
          int mid = (a + b) >> 1;
          if (__ldg(&pix[mid].col) < target)
            a = mid + 1;
          else
            b = mid;
        }
        return a;
      };
      const int a = lower(c), b = lower(c + W);
      if ((s << lr) >= r && (s << lr) + R <= r + W) {
        local += (unsigned long long)(b - a);
      } else {  // first / last strip: only the rows inside the window count
        for (int k = a; k < b; ++k) {
          const int row = (s << lr) + __ldg(&pix[k].q);
          local += (row >= r && row < r + W) ? 1 : 0;
        }
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    local += __shfl_down_sync(0xffffffffu, local, o);
    valid += __shfl_down_sync(0xffffffffu, valid, o);
  }
  if (lane == 0) {
    if (local) atomicAdd(out_nnz, local);
    if (valid) atomicAdd(out_valid, valid);
  }
}

// ------------------------------------------------------------------------------------------ expected by distance
// Per-diagonal sums of a region's stored upper-triangle pixels: what `cooltools expected-cis` feeds into the
// reference's --expected option (CLI.py:484-508; consumed at coolpup.py:861-918).  One warp per row.
__global__ void k_diag_sums(const int32_t* __restrict__ indptr_u, const int32_t* __restrict__ col_u,
                            const int32_t* __restrict__ cnt_u, const double* __restrict__ weight, int nb,
                            double* __restrict__ count_sum, double* __restrict__ bal_sum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < nb; r += nwarps) {
    const double wr = weight ? weight[r] : 1.0;
    for (int i = indptr_u[r] + lane; i < indptr_u[r + 1]; i += 32) {
      const int c = col_u[i];
      if (c < (int)r || c >= nb) continue;  // pixels that leave the region
      const double cnt = (double)cnt_u[i];
      atomicAdd(count_sum + (c - (int)r), cnt);  // raw counts: every stored pixel, masked bins included (cooltools)
      if (weight) {
        const double v = wr * weight[c] * cnt;
        if (v == v) atomicAdd(bal_sum + (c - (int)r), v);  // NaN: a masked bin, the position is not valid
      }
    }
  }
}

// n_valid[d] = number of positions (i, i + d) whose two bins are valid
//            = (nb - d) - #masked in [0, nb-d) - #masked in [d, nb) + #(i: i and i+d both masked)
__global__ void k_bad_pairs(const int32_t* __restrict__ badlist, int nbad, unsigned long long* __restrict__ both) {
  const int64_t total = (int64_t)nbad * nbad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i / nbad), b = (int)(i % nbad);
    if (a <= b) atomicAdd(both + (badlist[b] - badlist[a]), 1ull);
  }
}

__global__ void k_n_valid(const int32_t* __restrict__ badpre, const unsigned long long* __restrict__ both, int nb,
                          int64_t* __restrict__ n_valid) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= nb) return;
  int64_t v = nb - d;
  if (badpre != nullptr) {
    v -= badpre[nb - d];             // masked first bins
    v -= badpre[nb] - badpre[d];     // masked second bins
    v += (int64_t)both[d];           // counted twice
  }
  n_valid[d] = v;
}

// ------------------------------------------------------------------------------------------ control shifts (MT19937)
// The reference draws its random control shifts from numpy's global legacy RandomState (coolpup.py:392-396,
// 442-445): per block of n windows `np.random.randint(minshift, maxshift, n)` then `np.random.choice([-1, 1], n)`.
// numpy implements both on the MT19937 32-bit stream (numpy/random/src/mt19937/mt19937.c,
// distributions.c:random_bounded_uint64_fill -> buffered_bounded_masked_uint32): a randint is
// `low + v` for the first `v = next32() & mask` with `v <= rng` (mask = smallest 2^k - 1 >= rng = high - 1 - low);
// a choice is `next32() & 1`.  k_mt_shifts replays exactly that stream on the device, one CTA, so that a seeded run
// sees the same control windows as the reference without a host-side draw or a 1e7-row upload.
// Restated (and pinned against numpy) in oracle/mt19937_ref.py.
constexpr int MT_N = 624, MT_M = 397;
constexpr int MT_THREADS = 640;  // >= MT_N, a multiple of 32

struct pup_rng_state {
  uint32_t key[MT_N];
  int32_t pos;  // next word of key[] to use; MT_N = regenerate first
};

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// block-wide inclusive scan of one int per thread (MT_THREADS threads); returns the inclusive prefix, *total = sum
__device__ __forceinline__ int mt_block_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < MT_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < MT_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const int before = warp > 0 ? warp_sums[warp - 1] : 0;
  *total = warp_sums[MT_THREADS / 32 - 1];
  __syncthreads();  // warp_sums is reused by the next call
  return x + before;
}

// seg_n[s] randint draws followed by seg_n[s] sign draws per segment s; dbin[draw] = rint(shift * sign / resolution)
// in draw order (segment after segment); dbin == nullptr only advances the generator.
__global__ void __launch_bounds__(MT_THREADS) k_mt_shifts(pup_rng_state* st, const int64_t* __restrict__ seg_n,
                                                          int64_t n_seg, int64_t low, uint32_t rng, uint32_t mask,
                                                          double resolution, int32_t* __restrict__ dbin) {
  __shared__ uint32_t key[MT_N];
  __shared__ uint32_t out[MT_N];  // tempered words of the current block
  __shared__ int warp_sums[MT_THREADS / 32];
  __shared__ int s_cut;
  const int t = threadIdx.x;
  if (t < MT_N) key[t] = st->key[t];
  int pos = st->pos;
  __syncthreads();
  auto temper_all = [&]() {
    if (t < MT_N) {
      uint32_t y = key[t];
      y ^= y >> 11;
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= y >> 18;
      out[t] = y;
    }
    __syncthreads();
  };
  auto regenerate = [&]() {
    // key[i] = key[i + M] ^ twist(key[i], key[i + 1]): entries [0, N - M) read old values only, the next N - M read
    // entries the first phase wrote, and so on -- three dependent phases (+ the last word, which reads new key[0])
    uint32_t v = 0;
    if (t < MT_N - MT_M) v = mt_twist(key[t], key[t + 1], key[t + MT_M]);
    __syncthreads();
    if (t < MT_N - MT_M) key[t] = v;
    __syncthreads();
    if (t >= MT_N - MT_M && t < 2 * (MT_N - MT_M)) v = mt_twist(key[t], key[t + 1], key[t - (MT_N - MT_M)]);
    __syncthreads();
    if (t >= MT_N - MT_M && t < 2 * (MT_N - MT_M)) key[t] = v;
    __syncthreads();
    if (t >= 2 * (MT_N - MT_M) && t < MT_N - 1) v = mt_twist(key[t], key[t + 1], key[t - (MT_N - MT_M)]);
    __syncthreads();
    if (t >= 2 * (MT_N - MT_M) && t < MT_N - 1) key[t] = v;
    __syncthreads();
    if (t == MT_N - 1) key[t] = mt_twist(key[t], key[0], key[MT_M - 1]);
    __syncthreads();
    temper_all();
  };
  temper_all();  // (only meaningful for pos < MT_N)
  int64_t outbase = 0;
  for (int64_t seg = 0; seg < n_seg; ++seg) {
    const int64_t n = seg_n[seg];
    // ---- n accepted bounded integers
    int64_t got = 0;
    while (got < n) {
      if (pos >= MT_N) {
        regenerate();
        pos = 0;
      }
      const bool mine = t >= pos && t < MT_N;
      const uint32_t v = mine ? (out[t] & mask) : 0u;
      const int acc = (mine && v <= rng) ? 1 : 0;
      int total;
      const int incl = mt_block_scan(acc, warp_sums, &total);
      const int64_t need = n - got;
      if ((int64_t)total < need) {  // every word of the block is consumed (trailing rejects too: the draw goes on)
        if (acc && dbin) dbin[outbase + got + incl - 1] = (int32_t)(low + (int64_t)v);
        got += total;
        pos = MT_N;
      } else {  // the need-th accepted word ends this phase inside the block; words after it belong to the signs
        if (acc && (int64_t)incl <= need && dbin) dbin[outbase + got + incl - 1] = (int32_t)(low + (int64_t)v);
        if (acc && (int64_t)incl == need) s_cut = t;
        __syncthreads();
        pos = s_cut + 1;
        got = n;
        __syncthreads();
      }
    }
    // ---- n signs
    int64_t done = 0;
    while (done < n) {
      if (pos >= MT_N) {
        regenerate();
        pos = 0;
      }
      const int take = (int)min((int64_t)(MT_N - pos), n - done);
      if (t >= pos && t < pos + take && dbin) {
        const int64_t i = outbase + done + (t - pos);
        const int64_t sh = (int64_t)dbin[i] * ((out[t] & 1u) ? 1 : -1);
        dbin[i] = (int32_t)__double2ll_rn((double)sh / resolution);  // np.round: half to even
      }
      pos += take;
      done += take;
    }
    outbase += n;
    __syncthreads();
  }
  if (t < MT_N) st->key[t] = key[t];
  if (t == 0) st->pos = pos;
}

// ------------------------------------------------------------------------------------------ pair windows on device
// All-vs-all windows of one view region's bed features, in the reference's emission order (CoordCreator.
// get_combinations, coolpup.py:682-714: pairs (k, k + i) by offset i then k, distance-filtered; per offset block the
// ROI rows, then the nctrl shifted replicas, _control_regions 387-453) together with their accumulator slots.
struct PairGen {
  int m, nctrl, W, nb, nk, nf, targets, k_lo, k_hi, region_index;
  const int32_t* stbin;       // [m] region-relative first bin of every feature's window
  const double* center;       // [m]
  double mindist, maxdist;
  const int64_t* base;        // [m] kept pairs of all smaller offsets (exclusive prefix of per_offset)
  const int64_t* per_offset;  // [m]
  const int64_t* base_part;   // [m] the same two for the pairs whose row anchor k lies in [k_lo, k_hi) (the part written)
  const int64_t* per_offset_part;
  const int32_t* dbin;        // control shifts in draw order (k_mt_shifts) or null
  const int64_t* key1;        // [m] group-key part of a feature on side 1 / side 2 (null: 0)
  const int64_t* key2;
  const double* edges;        // distance-band edges (null: no band column)
  int n_edges;
  long long band_weight;
  int flip_mode;              // 0 none, 1: flipval[k] != 0, 2: flipval[k] > flipval[l]
  int swap_on_flip;           // ignore_group_order: a flipped window swaps the sides of its group key
  const int32_t* flipval;
  const int32_t* ident;       // by-window: feature identity (the window goes to both anchors' groups)
  int32_t *r0, *c0, *slot;    // outputs, [sum(per_offset_part) * (1 + nctrl) * targets]
  unsigned long long* first;  // [n_keys] atomicMin of (control?, region, emission position) over valid windows; null: off
  unsigned long long* n_roi;  // [1] valid ROI windows (x targets) of this call; null: off
};

__global__ void __launch_bounds__(256) k_pair_windows(const PairGen p) {
  __shared__ int warp_sums[2][8];
  __shared__ int s_base[2];
  const int i = blockIdx.x + 1;  // pair offset
  if (i >= p.m) return;
  const long long q = p.per_offset[i];
  const long long qp = p.per_offset_part[i];
  if (qp == 0) return;
  const long long blk = p.base[i] * (1 + p.nctrl);        // emission index of the block's first row
  const long long dr0 = p.base[i] * p.nctrl;              // first draw of the block
  const long long oblk = p.base_part[i] * (1 + p.nctrl);  // output index of the part's first row of this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
  __syncthreads();
  const int k_end = min(p.k_hi, p.m - i);  // anchors beyond the part contribute nothing (ranks only look backwards)
  for (int k0 = 0; k0 < k_end; k0 += blockDim.x) {
    const int k = k0 + threadIdx.x, l = k + i;
    bool keep = false;
    double dist = 0.0;
    if (l < p.m) {
      dist = p.center[l] - p.center[k];
      const double ad = fabs(dist);
      keep = p.mindist <= ad && ad <= p.maxdist;
    }
    const bool mine = keep && k >= p.k_lo && k < p.k_hi;
    // rank of this pair among the kept pairs of the offset (all anchors: emission position and draw index) and
    // among the kept pairs of the part (output index)
    const unsigned bal = __ballot_sync(0xffffffffu, keep), balp = __ballot_sync(0xffffffffu, mine);
    const unsigned below = (1u << lane) - 1u;
    if (lane == 0) {
      warp_sums[0][warp] = __popc(bal);
      warp_sums[1][warp] = __popc(balp);
    }
    __syncthreads();
    int before = s_base[0], beforep = s_base[1];
    for (int w = 0; w < warp; ++w) {
      before += warp_sums[0][w];
      beforep += warp_sums[1][w];
    }
    const long long j = before + __popc(bal & below);
    const long long jp = beforep + __popc(balp & below);
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0, totp = 0;
      for (int w = 0; w < 8; ++w) {
        tot += warp_sums[0][w];
        totp += warp_sums[1][w];
      }
      s_base[0] += tot;
      s_base[1] += totp;
    }
    __syncthreads();
    if (!mine) continue;
    // group key and flip flag are properties of the pair; the controls inherit them (coolpup.py:436-450)
    bool flip = false;
    if (p.flip_mode == 1) flip = p.flipval[k] != 0;
    if (p.flip_mode == 2) flip = p.flipval[k] > p.flipval[l];
    long long key = 0;
    if (p.ident == nullptr) {
      const bool sw = flip && p.swap_on_flip;
      if (p.key1) key += sw ? p.key1[l] : p.key1[k];
      if (p.key2) key += sw ? p.key2[k] : p.key2[l];
      if (p.edges) {  // np.searchsorted(edges, distance, side="right")
        int lo = 0, hi = p.n_edges;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (p.edges[mid] <= dist)
            lo = mid + 1;
          else
            hi = mid;
        }
        key += (long long)lo * p.band_weight;
      }
    }
    for (int rep = 0; rep <= p.nctrl; ++rep) {
      const long long pos = blk + (long long)rep * q + j;
      const int sh = rep == 0 ? 0 : p.dbin[dr0 + (long long)(rep - 1) * q + j];
      const int a = p.stbin[k] + sh, b = p.stbin[l] + sh;
      const bool valid = a >= 0 && b >= 0 && a + p.W <= p.nb && b + p.W <= p.nb;
      const int kind = rep == 0 ? 0 : 1;
      const long long o = (oblk + (long long)rep * qp + jp) * p.targets;
      for (int tg = 0; tg < p.targets; ++tg) {
        const long long kk = p.ident ? (long long)(tg == 0 ? p.ident[k] : p.ident[l]) : key;
        p.r0[o + tg] = a;
        p.c0[o + tg] = b;
        p.slot[o + tg] = (int32_t)((kk * p.nk + kind) * p.nf + (flip ? 1 : 0));
        if (valid && p.first) {
          const unsigned long long cand = ((unsigned long long)kind << 62) | ((unsigned long long)p.region_index << 40) |
                                          (unsigned long long)(pos * p.targets + tg);
          atomicMin(p.first + kk, cand);
        }
      }
      if (valid && kind == 0 && p.n_roi) atomicAdd(p.n_roi, (unsigned long long)p.targets);
    }
  }
}

// occ != nullptr: only query the occupancy; else launch
template <int R, int S, int PF, int MINB, bool QI>
cudaError_t launch_main_t(const MainParams& p, int grid, int threads, size_t smem, cudaStream_t st, int* occ) {
  auto kern = k_pileup_main<R, S, 4, PF, MINB, QI>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (occ) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, threads, smem);
  kern<<<grid, threads, smem, st>>>(p);
  return cudaGetLastError();
}

template <int R, int S>
cudaError_t launch_main_rs(const MainParams& p, int grid, int threads, size_t smem, cudaStream_t st, int* occ) {
  // tuning variants: PUP_TILE_INTERLEAVE=0 stores the tile rows of a strip one after the other instead of interleaved
  // column by column; PUP_PREFETCH=2 (default geometry only) adds an L2 prefetch two steps ahead (it paid while the
  // kernel still saw the dense windows: long runs; the sparse class left to it has short runs, -0.8 ms without;
  // wide windows, W >= 128, keep it)
  const int qi = env_int("PUP_TILE_INTERLEAVE", 1);
  if (R == 2 && S == 8 && env_int("PUP_PREFETCH", p.W >= 128 ? 2 : 0) == 0) return launch_main_t<R, S, 0, 2, true>(p, grid, threads, smem, st, occ);
  if (qi) return launch_main_t<R, S, 2, 2, true>(p, grid, threads, smem, st, occ);
  return launch_main_t<R, S, 2, 2, false>(p, grid, threads, smem, st, occ);
}

cudaError_t launch_main(int R, int S, const MainParams& p, int grid, int threads, size_t smem, cudaStream_t st,
                        int* occ) {
  if (R == 1 && S == 4) return launch_main_rs<1, 4>(p, grid, threads, smem, st, occ);
  if (R == 1 && S == 8) return launch_main_rs<1, 8>(p, grid, threads, smem, st, occ);
  if (R == 2 && S == 8) return launch_main_rs<2, 8>(p, grid, threads, smem, st, occ);
  if (R == 2 && S == 16) return launch_main_rs<2, 16>(p, grid, threads, smem, st, occ);
  if (R == 4 && S == 8) return launch_main_rs<4, 8>(p, grid, threads, smem, st, occ);
  if (R == 4 && S == 16) return launch_main_rs<4, 16>(p, grid, threads, smem, st, occ);
  if (R == 8 && S == 16) return launch_main_rs<8, 16>(p, grid, threads, smem, st, occ);
  if (R == 8 && S == 32) return launch_main_rs<8, 32>(p, grid, threads, smem, st, occ);
  return cudaErrorInvalidValue;
}

// strip geometry of a new region: R rows per strip, S lanes per strip run (PUP_STRIP / PUP_LANES override)
void choose_strip(int* R_out, int* S_out) {
  int R = env_int("PUP_STRIP", 2);
  if (R != 1 && R != 2 && R != 4 && R != 8) R = 2;
  int S = env_int("PUP_LANES", R == 1 ? 4 : R == 2 ? 8 : R == 4 ? 16 : 32);
  const bool ok = (R == 1 && (S == 4 || S == 8)) || (R == 2 && (S == 8 || S == 16)) ||
                  (R == 4 && (S == 8 || S == 16)) || (R == 8 && (S == 16 || S == 32));
  if (!ok) S = R == 1 ? 4 : R == 2 ? 8 : R == 4 ? 16 : 32;
  *R_out = R;
  *S_out = S;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

int pup_abi_version(void) { return 4; }

const char* pup_last_error(void) { return g_err.c_str(); }

int pup_last_launches(void) { return g_launches; }

int pup_device_count(int* n_out) {
  if (!n_out) return fail(PUP_E_ARG, "pup_device_count: null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *n_out = 0;
    return fail(PUP_E_NODEV, "cudaGetDeviceCount", e);
  }
  *n_out = n;
  return PUP_OK;
}

int pup_timing_enable(int on) {
  g_timing = on != 0;
  return PUP_OK;
}

int pup_timing_read(double* ms_by_tag, int* count_by_tag, int reset) {
  double ms[N_TAGS] = {0, 0, 0, 0, 0};
  int cnt[N_TAGS] = {0, 0, 0, 0, 0};
  for (auto& sp : g_spans) {
    float f = 0;
    cudaError_t e = cudaEventSynchronize(sp.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&f, sp.a, sp.b);
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "pup_timing_read", e);
    if (sp.tag >= 0 && sp.tag < N_TAGS) {
      ms[sp.tag] += f;
      cnt[sp.tag] += 1;
    }
  }
  if (reset) {
    for (auto& sp : g_spans) {
      cudaEventDestroy(sp.a);
      cudaEventDestroy(sp.b);
    }
    g_spans.clear();
  }
  for (int i = 0; i < N_TAGS; ++i) {
    if (ms_by_tag) ms_by_tag[i] = ms[i];
    if (count_by_tag) count_by_tag[i] = cnt[i];
  }
  return PUP_OK;
}

int64_t pup_acc_stride(int W) {
  if (W <= 0) return 0;
  return AccLayout(W).stride;
}

int64_t pup_region_device_bytes(const pup_region_t* r) { return r ? r->bytes : 0; }

}  // extern "C"

namespace {

// With the signed diagonal mask every pixel with col - row < ignore_diags is NaN in the reference's snippets: for
// ignore_diags >= 0 that is the whole lower triangle, which is then neither mirrored nor stored.
bool lower_triangle_masked(unsigned flags, int ignore_diags) { return !(flags & PUP_F_NODIAG) && ignore_diags >= 0; }

int choose_bucket_bits(int32_t nb, int32_t ns, int R, int64_t nnz, int* nbk_out) {
  // bucket width: aim at ~PUP_BUCKET_TARGET stored pixels per (strip, bucket); no table for very sparse rows
  double avg = (double)nnz / nb;  // pixels per row; a strip holds R times as many per column
  int target = env_int("PUP_BUCKET_TARGET", 8);
  if (avg <= 24.0) {
    *nbk_out = 1;
    return 31;
  }
  int lb = (int)floor(log2((double)target * nb / (avg * R)));
  if (lb < 2) lb = 2;
  while (((int64_t)((nb + (1 << lb) - 1) >> lb)) * ns * 4 > (int64_t)4 * nnz + (64 << 20) ||  // <= 25% of pixels
         ((int64_t)((nb + (1 << lb) - 1) >> lb)) * ns >= (1ll << 31))                           // 32-bit index
    ++lb;
  *nbk_out = (nb + (1 << lb) - 1) >> lb;
  return lb;
}

// The source matrix is on the device (row r = dcol/dcnt[rs[r] .. re[r]), r->nnz bounds the pixel count): normalise
// the pixels into the strip layout, build the bucket table and the masks.
int finish_region(pup_region* r, const int32_t* rs, const int32_t* re, const int32_t* dcol, const int32_t* dcnt,
                  const double* weight, int64_t nnz_estimate, Scratch& tmp) {
  // `weight` is DEVICE memory here (staged by the caller)
  cudaStream_t st = r->stream;
  const int32_t nb = r->nb;
  const int64_t nnz = r->nnz;
  choose_strip(&r->R, &r->S);
  r->lr = ilog2_ceil(r->R);
  r->ns = (nb + r->R - 1) >> r->lr;
  const int32_t ns = r->ns;
  const int S = r->S;
  r->lb = choose_bucket_bits(nb, ns, r->R, std::max<int64_t>(nnz_estimate, 1), &r->nbk);
  // strips padded with sentinel groups (+ slack for the main kernel's L2 prefetch two groups ahead)
  size_t n_ent = (size_t)nnz + (size_t)(2 * S) * ns + 4 * S + SENT_PAD * S;
  if (n_ent >= (1ull << 31)) return fail(PUP_E_ARG, "region create: padded pixel table exceeds 2^31 entries");
  CK(cudaMallocAsync((void**)&r->pix, n_ent * sizeof(Pix), st));
  CK(cudaMallocAsync((void**)&r->prow, (size_t)(ns + 1) * 4, st));
  {
    int32_t* plen;
    CK(tmp.alloc((void**)&plen, (size_t)(ns + 1) * 4));
    k_padded_len<<<(ns + 1 + 255) / 256, 256, 0, st>>>(rs, re, plen, nb, ns, r->lr, S);
    LAUNCH_CHECK("k_padded_len");
    size_t tb = 0;
    // pix[0 .. SENT_PAD * S) is a block of sentinel pixels (the main kernel reads it instead of predicating a load
    // off); the strips follow
    CK(cub::DeviceScan::ExclusiveScan(nullptr, tb, plen, r->prow, cub::Sum(), SENT_PAD * S, ns + 1, st));
    void* t;
    CK(tmp.alloc(&t, tb));
    CK(cub::DeviceScan::ExclusiveScan(t, tb, plen, r->prow, cub::Sum(), SENT_PAD * S, ns + 1, st));
    ++g_launches;
  }
  CK(cudaMallocAsync((void**)&r->bucket, (size_t)r->nbk * ns * 4, st));
  CK(cudaMallocAsync((void**)&r->ebad, (size_t)nb, st));
  CK(cudaMallocAsync((void**)&r->ebadpre, (size_t)(nb + 1) * 4, st));
  r->bytes += (int64_t)(n_ent * sizeof(Pix) + (size_t)(ns + 1) * 4 + (size_t)(nb + 1) * 4 + (size_t)r->nbk * ns * 4 +
                        (size_t)nb);
  const double* dw = weight;
  if (weight) {
    CK(cudaMallocAsync((void**)&r->bad, (size_t)nb, st));
    CK(cudaMallocAsync((void**)&r->badpre, (size_t)(nb + 1) * 4, st));
    CK(cudaMallocAsync((void**)&r->badlist, (size_t)nb * 4, st));
    r->bytes += (int64_t)nb * 9 + 4;
  }
  {
    int grid = std::min((ns + 7) / 8, 148 * 16);
    k_prepare_pixels<<<grid, 256, 0, st>>>(rs, re, r->prow, dcol, dcnt, dw, r->expected, r->pix, nb, ns, r->lr,
                                           r->ignore_diags, r->flags);
    LAUNCH_CHECK("k_prepare_pixels");
  }
  {
    int grid = std::min((ns + 7) / 8, 148 * 16);
    k_build_buckets<<<grid, 256, 0, st>>>(r->pix, r->prow, r->bucket, ns, r->nbk, r->lb, S);
    LAUNCH_CHECK("k_build_buckets");
  }
  {
    int32_t *ebad32, *bad32 = nullptr;
    CK(tmp.alloc((void**)&ebad32, (size_t)(nb + 1) * 4));
    if (weight) CK(tmp.alloc((void**)&bad32, (size_t)(nb + 1) * 4));
    k_masks<<<(nb + 1 + 255) / 256, 256, 0, st>>>(dw, r->expected, r->bad, r->ebad, ebad32, bad32, nb);
    LAUNCH_CHECK("k_masks");
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, ebad32, r->ebadpre, nb + 1, st));
    void* t;
    CK(tmp.alloc(&t, tb));
    CK(cub::DeviceScan::ExclusiveSum(t, tb, ebad32, r->ebadpre, nb + 1, st));
    ++g_launches;
    if (weight) {
      CK(cub::DeviceScan::ExclusiveSum(t, tb, bad32, r->badpre, nb + 1, st));
      k_badlist<<<(nb + 255) / 256, 256, 0, st>>>(r->bad, r->badpre, r->badlist, nb);
      ++g_launches;
      LAUNCH_CHECK("k_badlist");
    }
  }
  // dense diagonal band for k_pileup_dense: only when the lower triangle is masked (every cis pile-up), within a
  // memory budget of PUP_BAND_PCT percent of the pixel table (default 100: the region at most doubles in size);
  // the number of diagonals actually filled is decided on the device from the pixel density (no host round trip)
  // (matrices with fewer than 32 stored pixels per row cannot have a dense band worth building)
  if (env_int("PUP_BAND", 1) != 0 && lower_triangle_masked(r->flags, r->ignore_diags) && nnz >= 32ll * nb) {
    const int64_t budget = (int64_t)((double)(n_ent * sizeof(Pix)) * std::max(0, env_int("PUP_BAND_PCT", 100)) / 100.0);
    int64_t bw_max = std::min<int64_t>(budget / ((int64_t)nb * 8), nb);
    bw_max = std::min<int64_t>(bw_max, ((1ll << 31) - 1) / nb);  // k_pileup_dense indexes the band with 32 bits
    bw_max &= ~15ll;
    if (bw_max >= 64) {
      int hs = 5;
      while ((nb >> hs) >= BAND_HB) ++hs;
      int* hist;
      CK(tmp.alloc((void**)&hist, (size_t)BAND_HB * 4));
      CK(zero_async(hist, (size_t)BAND_HB * 4, st));
      CK(cudaMallocAsync((void**)&r->band_bw, 4, st));
      if (cudaMallocAsync((void**)&r->band, (size_t)nb * bw_max * 8, st) != cudaSuccess) {
        cudaGetLastError();  // no memory for the band: the region works without it (sparse kernel only)
        r->band = nullptr;
        return PUP_OK;
      }
      r->band_stride = (int)bw_max;
      r->bytes += (int64_t)nb * bw_max * 8 + 4;
      const int grid = std::min((ns + 7) / 8, 148 * 8);
      k_band_hist<<<grid, 256, 0, st>>>(r->pix, r->prow, ns, r->lr, hs, hist);
      LAUNCH_CHECK("k_band_hist");
      k_band_choose<<<1, 32, 0, st>>>(hist, nb, hs, (int)bw_max, std::max(1, env_int("PUP_BAND_DENSITY_PCT", 20)),
                                    std::max(0, r->ignore_diags), r->band_bw);
      LAUNCH_CHECK("k_band_choose");
      k_band_zero<<<148 * 8, 256, 0, st>>>(reinterpret_cast<double2*>(r->band), nb, r->band_stride, r->band_bw);
      LAUNCH_CHECK("k_band_zero");
      k_band_fill<<<grid, 256, 0, st>>>(r->pix, r->prow, ns, r->lr, r->band, r->band_stride, r->band_bw);
      LAUNCH_CHECK("k_band_fill");
    }
  }
  return PUP_OK;
}

int new_region(int device, int32_t nb, const double* expected, const double* coverage, int ignore_diags,
               unsigned flags, cudaStream_t st, pup_region** out) {
  pup_region* r = new pup_region();
  memset(r, 0, sizeof *r);
  r->device = device;
  r->nb = nb;
  r->stream = st;
  r->ignore_diags = ignore_diags;
  r->flags = flags;
  *out = r;
  // host vectors go through the internal copy stream like the matrix arrays (FIFO with them), not through the
  // caller's stream, where a small copy could wait behind uploads that other calls queued on the copy engine
  auto put = [&](double** dst, const double* src) -> int {
    r->bytes += (int64_t)nb * 8;
    cudaStream_t cs = is_device_ptr(src) ? st : copy_stream(device);
    if (!cs || cs == st) {
      CK(cudaMallocAsync((void**)dst, (size_t)nb * 8, st));
      CK(cudaMemcpyAsync(*dst, src, (size_t)nb * 8, cudaMemcpyDefault, st));
      return PUP_OK;
    }
    // allocated AND filled on the copy stream, like the matrix arrays (Uploader): an allocation on the caller's stream
    // would order this copy -- and every upload queued behind it -- after the previous region's preparation kernels
    CK(cudaMallocAsync((void**)dst, (size_t)nb * 8, cs));
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaMemcpyAsync(*dst, src, (size_t)nb * 8, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaEventRecord(ev, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "region create: vector upload", e);
    return PUP_OK;
  };
  if (expected) {
    int rc = put(&r->expected, expected);
    if (rc != PUP_OK) return rc;
  }
  if (coverage) {
    int rc = put(&r->coverage, coverage);
    if (rc != PUP_OK) return rc;
  }
  return PUP_OK;
}

int check_region_args(const char* who, int device, int32_t nb, int64_t nnz, const void* indptr, const void* col,
                      const void* count, const double* expected, unsigned flags) {
  if (nb <= 0 || nnz < 0 || nnz >= (1ll << 31) || !indptr || (nnz > 0 && (!col || !count)))
    return fail(PUP_E_ARG, "region create: bad sizes or null CSR arrays");
  if ((flags & PUP_F_OOE) && !expected) return fail(PUP_E_ARG, "region create: PUP_F_OOE needs an expected vector");
  if (flags & ~(PUP_F_OOE | PUP_F_NODIAG | PUP_F_ASYNC))
    return fail(PUP_E_ARG, "region create: only PUP_F_OOE / PUP_F_NODIAG / PUP_F_ASYNC apply");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(PUP_E_NODEV, "region create: no such CUDA device");
  }
  (void)who;
  return PUP_OK;
}

// Device copies of host input arrays: allocated and filled on the internal copy stream (so the upload does not
// queue behind the caller's stream), released on the caller's stream after use.
struct Uploader {
  cudaStream_t user, copy;
  std::vector<void*> ptrs;
  bool used = false;
  Uploader(cudaStream_t user_, int dev) : user(user_), copy(copy_stream(dev)) {
    if (!copy) copy = user;
  }
  ~Uploader() {
    for (void* p : ptrs) cudaFreeAsync(p, user);
  }
  // *dst = src if src is device memory, else a staged device copy
  int stage(const void* src, size_t bytes, const void** dst) {
    *dst = src;
    if (bytes == 0 || src == nullptr || is_device_ptr(src)) return PUP_OK;
    void* t;
    CK(cudaMallocAsync(&t, bytes, copy));
    ptrs.push_back(t);
    CK(cudaMemcpyAsync(t, src, bytes, cudaMemcpyHostToDevice, copy));
    *dst = t;
    used = true;
    return PUP_OK;
  }
  // make the caller's stream wait for all uploads issued so far
  int join() {
    if (!used || copy == user) return PUP_OK;
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, copy);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(user, ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "upload join", e);
    return PUP_OK;
  }
};

}  // namespace

extern "C" {

int pup_region_create(int device, int32_t nb, int64_t nnz, const int32_t* indptr, const int32_t* col,
                      const int32_t* count, const double* weight, const double* expected, const double* coverage,
                      int ignore_diags, unsigned flags, void* stream, pup_region_t** out) {
  if (!out) return fail(PUP_E_ARG, "pup_region_create: null output");
  *out = nullptr;
  int rc = check_region_args("pup_region_create", device, nb, nnz, indptr, col, count, expected, flags);
  if (rc != PUP_OK) return rc;
  const bool async = flags & PUP_F_ASYNC;
  flags &= ~PUP_F_ASYNC;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_region_create: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  g_launches = 0;
  pup_region* r = nullptr;
  rc = new_region(device, nb, expected, coverage, ignore_diags, flags, st, &r);
  if (rc == PUP_OK) {
    r->nnz = nnz;
    Scratch tmp(st);
    Uploader up(st, device);
    const void *dip = nullptr, *dcol = nullptr, *dcnt = nullptr, *dw = nullptr;
    rc = up.stage(indptr, (size_t)(nb + 1) * 4, &dip);
    if (rc == PUP_OK) rc = up.stage(col, (size_t)nnz * 4, &dcol);
    if (rc == PUP_OK) rc = up.stage(count, (size_t)nnz * 4, &dcnt);
    if (rc == PUP_OK) rc = up.stage(weight, (size_t)nb * 8, &dw);
    if (rc == PUP_OK) rc = up.join();
    if (rc == PUP_OK) {
      const int32_t *rs = (const int32_t*)dip, *re = (const int32_t*)dip + 1;
      int64_t nnz_est = nnz;
      if (lower_triangle_masked(flags, ignore_diags) && nnz > 0) {
        int32_t *trs, *tre;
        cudaError_t e = tmp.alloc((void**)&trs, (size_t)nb * 4);
        if (e == cudaSuccess) e = tmp.alloc((void**)&tre, (size_t)nb * 4);
        if (e != cudaSuccess) {
          rc = fail(PUP_E_OOM, "pup_region_create: scratch", e);
        } else {
          k_row_trim<<<(nb + 255) / 256, 256, 0, st>>>((const int32_t*)dip, (const int32_t*)dcol, nb, ignore_diags, trs,
                                                        tre);
          ++g_launches;
          rs = trs;
          re = tre;  // bucket sizing keeps using the symmetric count: the density per column is unchanged
        }
      }
      if (rc == PUP_OK)
        rc = finish_region(r, rs, re, (const int32_t*)dcol, (const int32_t*)dcnt, (const double*)dw, nnz_est, tmp);
    }
    cudaError_t e = cudaSuccess;
    const bool host_in = !is_device_ptr(indptr) || (nnz > 0 && (!is_device_ptr(col) || !is_device_ptr(count))) ||
                         (weight && !is_device_ptr(weight)) || (expected && !is_device_ptr(expected)) ||
                         (coverage && !is_device_ptr(coverage));
    if (rc == PUP_OK && host_in && !async) {
      e = cudaStreamSynchronize(st);  // host staging buffers must stay valid until the copies have been consumed
      if (e != cudaSuccess) rc = fail(PUP_E_CUDA, "pup_region_create: synchronize", e);
    }
  }
  if (rc != PUP_OK) {
    pup_region_destroy(r);
    return rc;
  }
  *out = r;
  return PUP_OK;
}

int pup_region_create_upper(int device, int32_t nb, int64_t nnz_upper, const int32_t* indptr_upper,
                            const int32_t* col_upper, const int32_t* count_upper, const double* weight,
                            const double* expected, const double* coverage, int ignore_diags, unsigned flags,
                            void* stream, pup_region_t** out) {
  if (!out) return fail(PUP_E_ARG, "pup_region_create_upper: null output");
  *out = nullptr;
  int rc = check_region_args("pup_region_create_upper", device, nb, nnz_upper, indptr_upper, col_upper, count_upper,
                             expected, flags);
  if (rc != PUP_OK) return rc;
  if (2 * nnz_upper >= (1ll << 31)) return fail(PUP_E_ARG, "pup_region_create_upper: symmetric matrix exceeds 2^31 pixels");
  const bool async = flags & PUP_F_ASYNC;
  flags &= ~PUP_F_ASYNC;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_region_create_upper: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  g_launches = 0;
  pup_region* r = nullptr;
  rc = new_region(device, nb, expected, coverage, ignore_diags, flags, st, &r);
  auto body = [&]() -> int {
    Scratch tmp(st);
    Uploader up(st, device);
    const void *v_ip, *v_col, *v_cnt, *v_w;
    int rc2;
    if ((rc2 = up.stage(indptr_upper, ((size_t)nb + 1) * 4, &v_ip)) != PUP_OK) return rc2;
    if ((rc2 = up.stage(col_upper, (size_t)nnz_upper * 4, &v_col)) != PUP_OK) return rc2;
    if ((rc2 = up.stage(count_upper, (size_t)nnz_upper * 4, &v_cnt)) != PUP_OK) return rc2;
    if ((rc2 = up.stage(weight, weight ? (size_t)nb * 8 : 0, &v_w)) != PUP_OK) return rc2;
    if ((rc2 = up.join()) != PUP_OK) return rc2;
    const int32_t *d_ip = (const int32_t*)v_ip, *d_col = (const int32_t*)v_col, *d_cnt = (const int32_t*)v_cnt;
    if (lower_triangle_masked(flags, ignore_diags)) {
      // nothing below the diagonal survives the mask: the uploaded upper triangle is all the pile-up needs
      int32_t *rs, *re;
      CK(tmp.alloc((void**)&rs, (size_t)nb * 4));
      CK(tmp.alloc((void**)&re, (size_t)nb * 4));
      k_row_trim<<<(nb + 255) / 256, 256, 0, st>>>(d_ip, d_col, nb, ignore_diags, rs, re);
      LAUNCH_CHECK("k_row_trim");
      r->nnz = nnz_upper;
      return finish_region(r, rs, re, d_col, d_cnt, (const double*)v_w, 2 * nnz_upper, tmp);
    }
    const size_t nu = (size_t)(nnz_upper > 0 ? nnz_upper : 1);
    int32_t *up_cnt, *lo_cnt, *lo_start, *tot_cnt, *key_a, *key_b, *val_a, *val_b, *row_of, *sym_indptr;
    CK(tmp.alloc((void**)&sym_indptr, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&up_cnt, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&lo_cnt, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&lo_start, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&tot_cnt, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&key_a, nu * 4));
    CK(tmp.alloc((void**)&key_b, nu * 4));
    CK(tmp.alloc((void**)&val_a, nu * 4));
    CK(tmp.alloc((void**)&val_b, nu * 4));
    CK(tmp.alloc((void**)&row_of, nu * 4));
    CK(zero_async(lo_cnt, (size_t)(nb + 1) * 4, st));
    CK(zero_async(up_cnt, (size_t)(nb + 1) * 4, st));
    const int wgrid = std::min((nb + 7) / 8, 148 * 16);
    k_upper_counts<<<wgrid, 256, 0, st>>>(d_ip, d_col, nb, up_cnt, lo_cnt, key_a, val_a);
    LAUNCH_CHECK("k_upper_counts");
    k_expand_rows<<<wgrid, 256, 0, st>>>(d_ip, row_of, nb);
    LAUNCH_CHECK("k_expand_rows");
    k_add_counts<<<(nb + 1 + 255) / 256, 256, 0, st>>>(up_cnt, lo_cnt, tot_cnt, nb);
    LAUNCH_CHECK("k_add_counts");
    size_t tb = 0, tb2 = 0;
    cub::DoubleBuffer<int32_t> dk(key_a, key_b), dv(val_a, val_b);
    const int bits = ilog2_ceil((int64_t)nb + 1);
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, tot_cnt, sym_indptr, nb + 1, st));
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int)nnz_upper, 0, bits, st));
    void* t;
    CK(tmp.alloc(&t, std::max(tb, tb2)));
    CK(cub::DeviceScan::ExclusiveSum(t, tb, tot_cnt, sym_indptr, nb + 1, st));
    CK(cub::DeviceScan::ExclusiveSum(t, tb, lo_cnt, lo_start, nb + 1, st));
    if (nnz_upper > 0) CK(cub::DeviceRadixSort::SortPairs(t, tb2, dk, dv, (int)nnz_upper, 0, bits, st));
    g_launches += 2 + (bits + 7) / 8;
    // the symmetric matrix has at most 2 * nnz_upper pixels: allocate for that bound instead of reading the exact
    // size back (no host synchronisation, so uploads can overlap the previous region's pile-up)
    r->nnz = 2 * nnz_upper;
    int32_t *col_s, *cnt_s;
    CK(tmp.alloc((void**)&col_s, (size_t)std::max<int64_t>(r->nnz, 1) * 4));
    CK(tmp.alloc((void**)&cnt_s, (size_t)std::max<int64_t>(r->nnz, 1) * 4));
    k_place_upper<<<wgrid, 256, 0, st>>>(d_ip, d_col, d_cnt, sym_indptr, lo_cnt, nb, col_s, cnt_s);
    LAUNCH_CHECK("k_place_upper");
    if (nnz_upper > 0) {
      k_place_lower<<<(unsigned)((nnz_upper + 255) / 256), 256, 0, st>>>(dk.Current(), dv.Current(), row_of, d_cnt,
                                                                          sym_indptr, lo_start, nb, col_s, cnt_s);
      LAUNCH_CHECK("k_place_lower");
    }
    return finish_region(r, sym_indptr, sym_indptr + 1, col_s, cnt_s, (const double*)v_w, r->nnz, tmp);
  };
  if (rc == PUP_OK) rc = body();
  const bool host_in = !is_device_ptr(indptr_upper) || (nnz_upper > 0 && (!is_device_ptr(col_upper) || !is_device_ptr(count_upper))) ||
                       (weight && !is_device_ptr(weight)) || (expected && !is_device_ptr(expected)) ||
                       (coverage && !is_device_ptr(coverage));
  if (rc == PUP_OK && host_in && !async) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(PUP_E_CUDA, "pup_region_create_upper: synchronize", e);
  }
  if (rc != PUP_OK) {
    pup_region_destroy(r);
    return rc;
  }
  *out = r;
  return PUP_OK;
}

int pup_upload(int device, void* dst, const void* src, int64_t bytes, void* stream) {
  if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(PUP_E_ARG, "pup_upload: bad arguments");
  if (bytes == 0) return PUP_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_upload: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  cudaStream_t cs = copy_stream(device);
  if (!cs) cs = st;
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  // the upload stream never waits for the caller's stream (that would hold back every upload queued behind this
  // one): dst must not be in use by work that is still pending when this is called
  cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess && cs != st) {
    e = cudaEventRecord(ev, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
  }
  cudaEventDestroy(ev);
  if (e != cudaSuccess) return fail(PUP_E_CUDA, "pup_upload", e);
  return PUP_OK;
}

int pup_region_destroy(pup_region_t* r) {
  if (!r) return PUP_OK;
  DeviceGuard guard(r->device);
  cudaStream_t st = r->stream;
  void* ptrs[] = {r->pix, r->prow, r->bucket,  r->expected, r->coverage, r->bad,
                  r->ebad, r->ebadpre, r->badpre, r->badlist, r->band, r->band_bw};
  for (void* p : ptrs)
    if (p) cudaFreeAsync(p, st);
  delete r;
  return PUP_OK;
}

int pup_accumulate(const pup_region_t* m, int64_t n_win, const int32_t* r0, const int32_t* c0, const int32_t* slot,
                   int W, int n_slots, unsigned flags, double* acc, void* stream, int64_t* n_valid_out) {
  if (!m) return fail(PUP_E_ARG, "pup_accumulate: null region");
  if (n_win < 0 || n_win >= (1ll << 31) || W <= 0 || n_slots <= 0 || !acc)
    return fail(PUP_E_ARG, "pup_accumulate: bad sizes or null accumulator");
  if (2 * W - 1 > VT) return fail(PUP_E_ARG, "pup_accumulate: W too large (max 256 bins)");
  if (n_win > 0 && (!r0 || !c0 || !slot)) return fail(PUP_E_ARG, "pup_accumulate: null window arrays");
  if (flags & ~(PUP_F_EXPCTRL | PUP_F_COVERAGE | PUP_F_ASYNC))
    return fail(PUP_E_ARG, "pup_accumulate: only PUP_F_EXPCTRL / PUP_F_COVERAGE / PUP_F_ASYNC apply (the others belong to the region)");
  const bool async = flags & PUP_F_ASYNC;
  flags &= ~PUP_F_ASYNC;
  if (async && n_valid_out) return fail(PUP_E_ARG, "pup_accumulate: PUP_F_ASYNC cannot return n_valid");
  if ((flags & PUP_F_EXPCTRL) && !m->expected)
    return fail(PUP_E_ARG, "pup_accumulate: expected requested but the region has none");
  if ((flags & PUP_F_EXPCTRL) && (m->flags & PUP_F_OOE))
    return fail(PUP_E_ARG, "pup_accumulate: PUP_F_EXPCTRL on a region prepared with PUP_F_OOE");
  if ((flags & PUP_F_COVERAGE) && !m->coverage)
    return fail(PUP_E_ARG, "pup_accumulate: coverage requested but the region has none");
  const int pb = ilog2_ceil((int64_t)m->nb + 1);
  const int lr = m->lr;
  // windows inside the region's dense band form a second class of every slot (k_window_keys), piled up by
  // k_pileup_dense
  const bool dense_ok = m->band != nullptr && env_int("PUP_DENSE", 1) != 0;
  const int n_cls = dense_ok ? 2 : 1;
  const int n_eslots0 = n_slots << lr;      // extended slots: (slot, r0 mod R)
  const int n_eslots = n_eslots0 * n_cls;  // ... of both classes
  const int sb = ilog2_ceil((int64_t)n_eslots + 1);
  if (2 * pb + sb > 62 || (int64_t)n_slots * n_cls << lr >= (1ll << 30))
    return fail(PUP_E_ARG, "pup_accumulate: too many slots for this region size");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_accumulate: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  g_launches = 0;
  if (n_valid_out) *n_valid_out = 0;
  if (n_win == 0) return PUP_OK;

  const unsigned all_flags = flags | m->flags;
  const AccLayout L(W);
  const int64_t acc_len = L.stride * n_slots;
  Scratch tmp(st);
  bool host_inputs = false;

  const int32_t *d_r0 = r0, *d_c0 = c0, *d_slot = slot;
  auto stage = [&](const int32_t*& d, const int32_t* h) -> cudaError_t {
    if (is_device_ptr(h)) return cudaSuccess;
    host_inputs = true;
    int32_t* t;
    cudaError_t e = tmp.alloc((void**)&t, (size_t)n_win * 4);
    if (e != cudaSuccess) return e;
    d = t;
    return cudaMemcpyAsync(t, h, (size_t)n_win * 4, cudaMemcpyHostToDevice, st);
  };
  CK(stage(d_r0, r0));
  CK(stage(d_c0, c0));
  CK(stage(d_slot, slot));

  double* d_acc = acc;
  const bool host_acc = !is_device_ptr(acc);
  if (host_acc && async) return fail(PUP_E_ARG, "pup_accumulate: PUP_F_ASYNC needs a device accumulator");
  if (host_acc) {
    CK(tmp.alloc((void**)&d_acc, (size_t)acc_len * 8));
    CK(zero_async(d_acc, (size_t)acc_len * 8, st));
  }

  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, m->device);
  // windows per work item: 64, fewer when the call is small (at least ~4 items per resident CTA, so that a launch of
  // a few thousand dense windows -- BASELINE configs[2] -- still keeps every SM busy)
  int ch = std::max(1, env_int("PUP_CHUNK", 64));
  if (env_int("PUP_CHUNK", 0) == 0) {
    const int64_t per_item = n_win / (int64_t)(n_sm * 2 * 4);
    ch = (int)std::max<int64_t>(4, std::min<int64_t>(64, per_item));
  }
  uint64_t *keys_a, *keys_b;
  int32_t *slot_start, *nchunks, *chunk_start;
  int* counters;  // [0] main work counter, [1] dense-num work counter, [2] slow windows of this call
  const uint64_t* keys;
  int2* win;
  {
    // 1. sort the windows by (slot, r0, c0) and cut every slot into chunks
    SpanGuard span(0, st);
    CK(tmp.alloc((void**)&keys_a, (size_t)n_win * 8));
    CK(tmp.alloc((void**)&keys_b, (size_t)n_win * 8));
    k_window_keys<<<(unsigned)((n_win + 255) / 256), 256, 0, st>>>(d_r0, d_c0, d_slot, keys_a, n_win, m->nb, W,
                                                                    n_slots, pb, lr, dense_ok ? m->band_bw : nullptr);
    LAUNCH_CHECK("k_window_keys");
    cub::DoubleBuffer<uint64_t> dbuf(keys_a, keys_b);
    size_t tb = 0;
    // valid keys use 2*pb+sb bits and have bit (2*pb+sb) clear; the invalid marker (~0) has it set,
    // so sorting bits [0, 2*pb+sb+1) orders everything and puts the invalid windows last.
    const int end_bit = 2 * pb + sb + 1;
    // the order of the lowest c0 bits only matters for locality: leave up to 4 of them unsorted when that saves
    // an 8-bit radix pass (the key still carries them for k_decode_windows)
    int begin_bit = end_bit - 8 * ((end_bit + 7) / 8 - 1);
    if (begin_bit > 4 || begin_bit >= pb || end_bit <= 8) begin_bit = 0;
    // Only the PUP_SORT_C0_BITS (default 4) most significant c0 bits are sorted: windows of one row r0 then stay
    // within 1/16 of the chromosome of each other, which is all the L2 locality the pile-up needs (measured: same
    // main-kernel time as a full sort, 0.4 ms less sorting per 1.1e7 windows); -1 = sort every bit
    const int c0_bits = env_int("PUP_SORT_C0_BITS", 4);
    if (c0_bits >= 0 && c0_bits < pb) begin_bit = pb - c0_bits;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dbuf, (int)n_win, begin_bit, end_bit, st));
    void* t;
    CK(tmp.alloc(&t, tb));
    CK(cub::DeviceRadixSort::SortKeys(t, tb, dbuf, (int)n_win, begin_bit, end_bit, st));
    g_launches += (end_bit - begin_bit + 7) / 8 + 1;
    keys = dbuf.Current();
    CK(tmp.alloc((void**)&win, (size_t)n_win * 8));
    k_decode_windows<<<(unsigned)((n_win + 255) / 256), 256, 0, st>>>(keys, win, n_win, pb);
    LAUNCH_CHECK("k_decode_windows");

    CK(tmp.alloc((void**)&slot_start, (size_t)(n_eslots + 1) * 4));
    CK(tmp.alloc((void**)&nchunks, (size_t)(n_eslots + 1) * 4));
    CK(tmp.alloc((void**)&chunk_start, (size_t)(n_eslots + 1) * 4));
    CK(tmp.alloc((void**)&counters, 16));
    CK(zero_async(counters, 16, st));
    k_slot_bounds<<<(n_eslots + 1 + 127) / 128, 128, 0, st>>>(keys, (int)n_win, n_eslots, pb, ch, slot_start,
                                                               nchunks);
    LAUNCH_CHECK("k_slot_bounds");
    size_t tb2 = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb2, nchunks, chunk_start, n_eslots + 1, st));
    void* t2;
    CK(tmp.alloc(&t2, tb2));
    CK(cub::DeviceScan::ExclusiveSum(t2, tb2, nchunks, chunk_start, n_eslots + 1, st));
    ++g_launches;
  }
  WinCtx ctx{m->nb, W, pb, lr, m->ignore_diags, all_flags, m->ebadpre, n_slots};
  ChunkTable chunks{slot_start, chunk_start, n_eslots, ch};          // every window (count / vector / dense-num kernels)
  ChunkTable chunks_sparse{slot_start, chunk_start, n_eslots0, ch};  // the class k_pileup_main piles up

  // 2. per-window counts (n, n_fast, masked rows / columns) and, on request, the O(W) fp64 vectors
  {
    SpanGuard span(1, st);
    const int64_t cstride = (int64_t)W * W + 2 * W + 2;
    const int64_t one = (int64_t)n_slots * cstride;
    const int copies = (int)std::max<int64_t>(1, std::min<int64_t>(64, (64ll << 20) / (one * 4)));
    int* counts;
    CK(tmp.alloc((void**)&counts, (size_t)(one * copies) * 4));
    CK(zero_async(counts, (size_t)(one * copies) * 4, st));
    CountParams cp{ctx, keys, slot_start, n_slots, n_eslots, m->badpre, m->badlist, counts, copies, counters + 2};
    const size_t csmem = (size_t)cstride * 4;
    // the shared-memory variant pays a scan of the tile per slot change: only for small tiles (W <= 108) and long slot
    // runs (configs[4], W = 203 with 32 slots: 1.8 -> 3.2 ms with it)
    if (csmem <= 48 * 1024 && n_win >= 1024 * (int64_t)n_slots && env_int("PUP_COUNTS_SMEM", 1) != 0) {
      CK(cudaFuncSetAttribute(k_window_counts_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      int cocc = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cocc, k_window_counts_smem, 256, csmem));
      const int cgrid = (int)std::min<int64_t>((n_win + 1023) / 1024, (int64_t)n_sm * std::max(1, cocc));
      k_window_counts_smem<<<cgrid, 256, csmem, st>>>(cp);
      LAUNCH_CHECK("k_window_counts_smem");
    } else {
      k_window_counts<<<(unsigned)((n_win + 255) / 256), 256, 0, st>>>(cp);
      LAUNCH_CHECK("k_window_counts");
    }
    int rgrid = (int)std::min<int64_t>((one + 255) / 256, (int64_t)n_sm * 8);
    k_counts_reduce<<<rgrid, 256, 0, st>>>(counts, copies, n_slots, W, lr, n_cls, slot_start, d_acc);
    LAUNCH_CHECK("k_counts_reduce");
    if (flags & (PUP_F_EXPCTRL | PUP_F_COVERAGE)) {
      VecParams vp{ctx, keys, slot_start, n_eslots, m->expected, m->coverage, d_acc};
      const int need = ((flags & PUP_F_EXPCTRL) ? 2 * W - 1 : W);
      const int vthreads = std::min(VT, ((need + 31) / 32) * 32);
      int grid = (int)std::min<int64_t>((n_win + VCH - 1) / VCH, (int64_t)n_sm * 16);
      k_vector<<<grid, vthreads, 0, st>>>(vp);
      LAUNCH_CHECK("k_vector");
    }
  }

  // 3. the pile-up itself
  {
    SpanGuard span(2, st);
    const int R = m->R, S = m->S;
    // one lane group per strip of the window; a band = the groups whose R x W fp64 tile rows fit PUP_TILE_KB
    const int n_groups = (W + 2 * R - 2) / R;  // ceil((W + R - 1) / R): a window touches at most that many strips
    // tile budget per CTA: 112 KB leaves two CTAs per SM; W = 203 then takes 3 row bands of 34 strips (measured on
    // configs[4]: 12.4 ms per step against 18.1 ms with 72 KB / 5 bands and 20.9 ms with 96 KB)
    const int tile_kb = env_int("PUP_TILE_KB", 112);
    // tile row stride: W plus PUP_TILE_PAD doubles (bank mapping of the rows a half-warp works on)
    const int TW = W + std::max(0, env_int("PUP_TILE_PAD", 0));
    int Gb = (int)std::min<int64_t>(n_groups, ((int64_t)tile_kb * 1024) / (8ll * TW * R));
    const int nt_max = (S == 32) ? NT_MAX : NT_MAX - 32;  // the kernel's launch bound
    Gb = std::min(Gb, nt_max / S);
    if (Gb < 1) Gb = 1;
    const int n_bands = (n_groups + Gb - 1) / Gb;
    Gb = (n_groups + n_bands - 1) / n_bands;  // balance the bands
    const int threads = std::min(nt_max, ((Gb * S + 31) / 32) * 32);
    const bool dynamic = env_int("PUP_SCHED", 1) != 0;
    const size_t smem = (size_t)Gb * R * TW * 8 + sizeof(DynSched);
    MainParams mp{W, m->ns, m->lb, m->pix, m->bucket, win, chunks_sparse, Gb, n_groups, n_bands, TW,
                  dynamic ? counters : nullptr, d_acc};
    int occ = 1;
    cudaError_t e = launch_main(R, S, mp, 0, threads, smem, st, &occ);
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "main kernel occupancy query", e);
    if (occ < 1) return fail(PUP_E_CUDA, "main kernel does not fit on an SM");
    e = launch_main(R, S, mp, n_sm * occ, threads, smem, st, nullptr);
    ++g_launches;
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "launch k_pileup_main", e);
    span.close();
    if (dense_ok) {
      SpanGuard span_dense(4, st);
      // thread t owns tile cells t, t + T, ... (of its row band, when the tile is cut): 256 threads (4 CTAs per SM)
      // up to 1024 cells, else 1024 threads with up to 7 cells each
      const int w2 = W * W;
      const int n_bands_d = (w2 + DENSE_BAND_CELLS - 1) / DENSE_BAND_CELLS;
      const int rows_pb = (W + n_bands_d - 1) / n_bands_d;
      const int cells_pb = rows_pb * W;
      DenseParams dp{W, m->band_stride, lr, n_slots, rows_pb, (W + rows_pb - 1) / rows_pb, m->band, win, chunks, n_eslots0,
                     counters + 3, d_acc};
      if (cells_pb <= 256 * 2)
        k_pileup_dense<256, 2><<<n_sm * 4, 256, 0, st>>>(dp);
      else if (cells_pb <= 256 * 4)
        k_pileup_dense<256, 4><<<n_sm * 4, 256, 0, st>>>(dp);
      else if (cells_pb <= 1024 * 2)
        k_pileup_dense<1024, 2><<<n_sm, 1024, 0, st>>>(dp);
      else if (cells_pb <= 1024 * 4)
        k_pileup_dense<1024, 4><<<n_sm, 1024, 0, st>>>(dp);
      else if (cells_pb <= 1024 * 7)
        k_pileup_dense<1024, 7><<<n_sm, 1024, 0, st>>>(dp);
      else
        k_pileup_dense<1024, 8><<<n_sm, 1024, 0, st>>>(dp);  // rows_pb * W can exceed the band budget by < W cells
      LAUNCH_CHECK("k_pileup_dense");
    }
  }

  // 4. dense pixel counts of the slow windows (returns at once when the call has none)
  {
    SpanGuard span(3, st);
    const size_t smem = (size_t)W * W * 4;
    if (smem > 220 * 1024) return fail(PUP_E_ARG, "pup_accumulate: W too large for the dense-num tile");
    SlowParams sp{ctx, m->bad, m->ebad, keys, chunks, d_acc, counters + 2, counters + 1};
    CK(cudaFuncSetAttribute(k_num_slow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_num_slow, 512, smem));
    if (occ < 1) return fail(PUP_E_CUDA, "dense-num kernel does not fit on an SM");
    k_num_slow<<<n_sm * occ, 512, smem, st>>>(sp);
    LAUNCH_CHECK("k_num_slow");
  }

  // 5. outputs that live on the host
  if (host_acc) {
    std::vector<double> h((size_t)acc_len);
    CK(cudaMemcpyAsync(h.data(), d_acc, (size_t)acc_len * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < acc_len; ++i) acc[i] += h[(size_t)i];
  }
  if (n_valid_out) {
    int32_t nv = 0;
    CK(cudaMemcpyAsync(&nv, slot_start + n_eslots, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_valid_out = nv;
  } else if (host_inputs && !host_acc && !async) {
    CK(cudaStreamSynchronize(st));
  }
  return PUP_OK;
}

int pup_accumulate_region(int device, int32_t nb, int64_t nnz, const int32_t* indptr, const int32_t* col,
                          const int32_t* count, const double* weight, const double* expected,
                          const double* coverage, int64_t n_win, const int32_t* r0, const int32_t* c0,
                          const int32_t* slot, int W, int ignore_diags, int n_slots, unsigned flags, double* acc,
                          void* stream, int64_t* n_valid_out) {
  pup_region_t* r = nullptr;
  if (flags & PUP_F_ASYNC) return fail(PUP_E_ARG, "pup_accumulate_region: PUP_F_ASYNC needs the two-call form");
  int rc = pup_region_create(device, nb, nnz, indptr, col, count, weight, expected, coverage, ignore_diags,
                             flags & (PUP_F_OOE | PUP_F_NODIAG), stream, &r);
  if (rc != PUP_OK) return rc;
  int l0 = g_launches;
  rc = pup_accumulate(r, n_win, r0, c0, slot, W, n_slots, flags & (PUP_F_EXPCTRL | PUP_F_COVERAGE), acc, stream,
                      n_valid_out);
  g_launches += l0;
  pup_region_destroy(r);
  return rc;
}

int pup_acc_export(const double* acc, int W, int n_slots, int device, void* stream, double* sum, int64_t* num,
                   int64_t* n, double* cov_start, double* cov_end, double* exp_sum, int64_t* exp_num) {
  if (!acc || W <= 0 || n_slots <= 0) return fail(PUP_E_ARG, "pup_acc_export: bad arguments");
  const AccLayout L(W);
  const int64_t len = L.stride * n_slots;
  std::vector<double> hbuf;
  const double* h = acc;
  if (is_device_ptr(acc)) {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(PUP_E_NODEV, "pup_acc_export: cudaSetDevice failed");
    hbuf.resize((size_t)len);
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(hbuf.data(), acc, (size_t)len * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h = hbuf.data();
  }
  for (int s = 0; s < n_slots; ++s) {
    const double* a = h + (int64_t)s * L.stride;
    const double nfast = a[L.off_nfast];
    if (n) n[s] = (int64_t)llround(a[L.off_n]);
    for (int i = 0; i < W; ++i) {
      if (cov_start) cov_start[(int64_t)s * W + i] = a[L.off_covs + i];
      if (cov_end) cov_end[(int64_t)s * W + i] = a[L.off_cove + i];
      for (int j = 0; j < W; ++j) {
        const int64_t o = (int64_t)s * L.w2 + (int64_t)i * W + j;
        if (sum) sum[o] = a[(int64_t)i * W + j];
        if (num) num[o] = (int64_t)llround(nfast - a[L.off_rb + i] - a[L.off_cb + j] + a[L.off_num + (int64_t)i * W + j]);
        if (exp_sum) exp_sum[o] = a[L.off_tsum + (j - i + W - 1)];
        if (exp_num) exp_num[o] = (int64_t)llround(a[L.off_tnum + (j - i + W - 1)]);
      }
    }
  }
  return PUP_OK;
}

int pup_stripes(const pup_region_t* m, int64_t n_win, const int32_t* r0, const int32_t* c0, int W, double* horizontal,
                double* vertical, void* stream) {
  if (!m || n_win < 0 || W <= 0 || !horizontal || !vertical) return fail(PUP_E_ARG, "pup_stripes: bad arguments");
  if (n_win == 0) return PUP_OK;
  if (!r0 || !c0) return fail(PUP_E_ARG, "pup_stripes: null window arrays");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_stripes: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch tmp(st);
  const int32_t *d_r0 = r0, *d_c0 = c0;
  if (!is_device_ptr(r0)) {
    int32_t* t;
    CK(tmp.alloc((void**)&t, (size_t)n_win * 4));
    CK(cudaMemcpyAsync(t, r0, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
    d_r0 = t;
  }
  if (!is_device_ptr(c0)) {
    int32_t* t;
    CK(tmp.alloc((void**)&t, (size_t)n_win * 4));
    CK(cudaMemcpyAsync(t, c0, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
    d_c0 = t;
  }
  const size_t bytes = (size_t)n_win * W * 8;
  double *d_h = horizontal, *d_v = vertical;
  const bool host_h = !is_device_ptr(horizontal), host_v = !is_device_ptr(vertical);
  if (host_h) CK(tmp.alloc((void**)&d_h, bytes));
  if (host_v) CK(tmp.alloc((void**)&d_v, bytes));
  StripeParams sp{m->pix, m->prow, m->bad, (m->flags & PUP_F_OOE) ? m->expected : nullptr, m->nb, W,
                  m->ignore_diags, m->lr, m->flags};
  k_stripes<<<(unsigned)((n_win * 32 + 255) / 256), 256, 0, st>>>(sp, d_r0, d_c0, n_win, d_h, d_v);
  LAUNCH_CHECK("k_stripes");
  if (host_h) CK(cudaMemcpyAsync(horizontal, d_h, bytes, cudaMemcpyDeviceToHost, st));
  if (host_v) CK(cudaMemcpyAsync(vertical, d_v, bytes, cudaMemcpyDeviceToHost, st));
  if (host_h || host_v || !is_device_ptr(r0) || !is_device_ptr(c0)) CK(cudaStreamSynchronize(st));
  return PUP_OK;
}

int pup_expected_cis(int device, int32_t nb, int64_t nnz_upper, const int32_t* indptr_upper, const int32_t* col_upper,
                     const int32_t* count_upper, const double* weight, double* count_sum, double* balanced_sum,
                     int64_t* n_valid, void* stream) {
  if (nb <= 0 || nnz_upper < 0 || nnz_upper >= (1ll << 31) || !indptr_upper || (nnz_upper > 0 && (!col_upper || !count_upper)) ||
      !count_sum || !n_valid)
    return fail(PUP_E_ARG, "pup_expected_cis: bad sizes or null arrays");
  if ((weight == nullptr) != (balanced_sum == nullptr))
    return fail(PUP_E_ARG, "pup_expected_cis: weight and balanced_sum go together");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return fail(PUP_E_NODEV, "pup_expected_cis: no such CUDA device");
  }
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_expected_cis: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  g_launches = 0;
  Scratch tmp(st);
  Uploader up(st, device);
  const void *v_ip, *v_col, *v_cnt, *v_w;
  int rc;
  if ((rc = up.stage(indptr_upper, ((size_t)nb + 1) * 4, &v_ip)) != PUP_OK) return rc;
  if ((rc = up.stage(col_upper, (size_t)nnz_upper * 4, &v_col)) != PUP_OK) return rc;
  if ((rc = up.stage(count_upper, (size_t)nnz_upper * 4, &v_cnt)) != PUP_OK) return rc;
  if ((rc = up.stage(weight, weight ? (size_t)nb * 8 : 0, &v_w)) != PUP_OK) return rc;
  if ((rc = up.join()) != PUP_OK) return rc;
  const double* dw = (const double*)v_w;
  double *d_cs = count_sum, *d_bs = balanced_sum;
  int64_t* d_nv = n_valid;
  const bool h_cs = !is_device_ptr(count_sum), h_bs = balanced_sum && !is_device_ptr(balanced_sum),
             h_nv = !is_device_ptr(n_valid);
  if (h_cs) CK(tmp.alloc((void**)&d_cs, (size_t)nb * 8));
  if (h_bs) CK(tmp.alloc((void**)&d_bs, (size_t)nb * 8));
  if (h_nv) CK(tmp.alloc((void**)&d_nv, (size_t)nb * 8));
  CK(zero_async(d_cs, (size_t)nb * 8, st));
  if (d_bs) CK(zero_async(d_bs, (size_t)nb * 8, st));
  const int wgrid = std::min((nb + 7) / 8, 148 * 16);
  k_diag_sums<<<wgrid, 256, 0, st>>>((const int32_t*)v_ip, (const int32_t*)v_col, (const int32_t*)v_cnt, dw, nb, d_cs,
                                     d_bs);
  LAUNCH_CHECK("k_diag_sums");
  int32_t* badpre = nullptr;
  unsigned long long* both = nullptr;
  if (dw) {
    uint8_t *bad, *ebad;
    int32_t *ebad32, *bad32, *badlist;
    CK(tmp.alloc((void**)&bad, (size_t)nb));
    CK(tmp.alloc((void**)&ebad, (size_t)nb));
    CK(tmp.alloc((void**)&ebad32, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&bad32, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&badpre, (size_t)(nb + 1) * 4));
    CK(tmp.alloc((void**)&badlist, (size_t)nb * 4));
    CK(tmp.alloc((void**)&both, (size_t)nb * 8));
    CK(zero_async(both, (size_t)nb * 8, st));
    k_masks<<<(nb + 1 + 255) / 256, 256, 0, st>>>(dw, nullptr, bad, ebad, ebad32, bad32, nb);
    LAUNCH_CHECK("k_masks");
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, bad32, badpre, nb + 1, st));
    void* t;
    CK(tmp.alloc(&t, tb));
    CK(cub::DeviceScan::ExclusiveSum(t, tb, bad32, badpre, nb + 1, st));
    ++g_launches;
    k_badlist<<<(nb + 255) / 256, 256, 0, st>>>(bad, badpre, badlist, nb);
    LAUNCH_CHECK("k_badlist");
    int32_t nbad = 0;  // the pair kernel's grid does not depend on it, only its loop bound: read it on the device
    CK(cudaMemcpyAsync(&nbad, badpre + nb, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (nbad > 0) {
      k_bad_pairs<<<148 * 8, 256, 0, st>>>(badlist, nbad, both);
      LAUNCH_CHECK("k_bad_pairs");
    }
  }
  k_n_valid<<<(nb + 255) / 256, 256, 0, st>>>(badpre, both, nb, d_nv);
  LAUNCH_CHECK("k_n_valid");
  if (h_cs) CK(cudaMemcpyAsync(count_sum, d_cs, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
  if (h_bs) CK(cudaMemcpyAsync(balanced_sum, d_bs, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
  if (h_nv) CK(cudaMemcpyAsync(n_valid, d_nv, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
  const bool host_in = !is_device_ptr(indptr_upper) || (nnz_upper > 0 && (!is_device_ptr(col_upper) || !is_device_ptr(count_upper))) ||
                       (weight && !is_device_ptr(weight));
  if (h_cs || h_bs || h_nv || host_in) CK(cudaStreamSynchronize(st));
  return PUP_OK;
}

// ------------------------------------------------------------------------------------------ host-side window layout
// (plain host code: no device involved)
int64_t pup_pair_windows_count(int32_t m, const double* center, double mindist, double maxdist, int64_t* per_offset) {
  return pup_pair_windows_count_range(m, center, mindist, maxdist, 0, m, per_offset);
}

int64_t pup_pair_windows_count_range(int32_t m, const double* center, double mindist, double maxdist, int32_t k_lo,
                                     int32_t k_hi, int64_t* per_offset) {
  if (m < 0 || (m > 0 && (!center || !per_offset)) || k_lo < 0 || k_hi < k_lo || k_hi > m) {
    fail(PUP_E_ARG, "pup_pair_windows_count: bad arguments");
    return -1;
  }
  int64_t total = 0;
  if (m > 0) per_offset[0] = 0;
  for (int32_t i = 1; i < m; ++i) {
    int64_t q = 0;
    for (int32_t k = k_lo; k < k_hi && k + i < m; ++k) {
      const double d = std::fabs(center[k + i] - center[k]);
      q += (mindist <= d && d <= maxdist) ? 1 : 0;
    }
    per_offset[i] = q;
    total += q;
  }
  return total;
}

int pup_pair_windows_fill(int32_t m, const int64_t* stbin, const double* center, double mindist, double maxdist,
                          int32_t nctrl, const int64_t* dbin, int64_t* st1, int64_t* st2, int8_t* kind, int64_t* idx1,
                          int64_t* idx2, double* distance) {
  if (m < 0 || nctrl < 0 || (m > 1 && (!stbin || !center || !st1 || !st2 || !kind || !idx1 || !idx2 || !distance)))
    return fail(PUP_E_ARG, "pup_pair_windows_fill: bad arguments");
  int64_t out = 0, drawn = 0;
  std::vector<int32_t> ks;
  for (int32_t i = 1; i < m; ++i) {
    ks.clear();
    for (int32_t k = 0; k + i < m; ++k) {
      const double d = std::fabs(center[k + i] - center[k]);
      if (mindist <= d && d <= maxdist) ks.push_back(k);
    }
    const int64_t n = (int64_t)ks.size();
    if (n == 0) continue;
    if (nctrl > 0 && !dbin) return fail(PUP_E_ARG, "pup_pair_windows_fill: control shifts missing");
    for (int32_t rep = 0; rep <= nctrl; ++rep) {  // rep 0: the ROI rows of the block, then its nctrl shifted replicas
      for (int64_t j = 0; j < n; ++j) {
        const int32_t k = ks[(size_t)j], l = k + i;
        const int64_t sh = rep == 0 ? 0 : dbin[drawn + (int64_t)(rep - 1) * n + j];
        st1[out] = stbin[k] + sh;
        st2[out] = stbin[l] + sh;
        kind[out] = rep == 0 ? 0 : 1;
        idx1[out] = k;
        idx2[out] = l;
        distance[out] = center[l] - center[k];
        ++out;
      }
    }
    drawn += n * nctrl;
  }
  return PUP_OK;
}

int pup_accumulate_rescaled(const pup_region_t* m, int64_t n_win, const int32_t* r0, const int32_t* c0,
                            const int32_t* h, const int32_t* w, const int32_t* slot, const int32_t* mode, int rescale_size,
                            int n_slots, unsigned flags, double* acc, void* stream, int64_t* n_valid_out) {
  if (!m) return fail(PUP_E_ARG, "pup_accumulate_rescaled: null region");
  if (n_win < 0 || n_win >= (1ll << 31) || rescale_size <= 0 || rescale_size > 255 || n_slots <= 0 || !acc)
    return fail(PUP_E_ARG, "pup_accumulate_rescaled: bad sizes or null accumulator");
  if (n_win > 0 && (!r0 || !c0 || !h || !w || !slot)) return fail(PUP_E_ARG, "pup_accumulate_rescaled: null window arrays");
  if (flags & ~(PUP_F_COVERAGE | PUP_F_LOCAL | PUP_F_ASYNC))
    return fail(PUP_E_ARG, "pup_accumulate_rescaled: only PUP_F_COVERAGE / PUP_F_LOCAL / PUP_F_ASYNC apply");
  const bool async = flags & PUP_F_ASYNC;
  flags &= ~PUP_F_ASYNC;
  if (async && n_valid_out) return fail(PUP_E_ARG, "pup_accumulate_rescaled: PUP_F_ASYNC cannot return n_valid");
  if ((flags & PUP_F_COVERAGE) && !m->coverage) return fail(PUP_E_ARG, "pup_accumulate_rescaled: the region has no coverage");
  if (mode && !m->expected) return fail(PUP_E_ARG, "pup_accumulate_rescaled: expected-block windows need a region with expected");
  if (!is_device_ptr(acc)) return fail(PUP_E_ARG, "pup_accumulate_rescaled: the accumulator must be device memory");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_accumulate_rescaled: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  g_launches = 0;
  if (n_valid_out) *n_valid_out = 0;
  if (n_win == 0) return PUP_OK;
  Scratch tmp(st);
  // the window arrays are small next to the snippets: staged from host memory when needed; their maxima size the
  // per-CTA scratch, so host copies are required (device arrays are read back)
  std::vector<int32_t> hh((size_t)n_win), hw((size_t)n_win);
  auto to_host = [&](const int32_t* src, int32_t* dst) -> cudaError_t {
    if (!is_device_ptr(src)) {
      memcpy(dst, src, (size_t)n_win * 4);
      return cudaSuccess;
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)n_win * 4, cudaMemcpyDeviceToHost, st);
    return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
  };
  CK(to_host(h, hh.data()));
  CK(to_host(w, hw.data()));
  int64_t max_cells = 1;
  int max_side = 1;
  for (int64_t i = 0; i < n_win; ++i) {
    if (hh[(size_t)i] < 0 || hw[(size_t)i] < 0) continue;
    max_cells = std::max<int64_t>(max_cells, (int64_t)hh[(size_t)i] * hw[(size_t)i]);
    max_side = std::max(max_side, std::max(hh[(size_t)i], hw[(size_t)i]));
  }
  const int rs = rescale_size;
  const int max_tmp = rs * std::max(1, (max_side + rs - 1) / rs);
  const size_t smem = (((size_t)rs * rs * 12 + 15) / 16) * 16 + 2 * (size_t)max_tmp * sizeof(ZoomPlan);
  if (smem > 220 * 1024) return fail(PUP_E_ARG, "pup_accumulate_rescaled: windows too large for the zoom plans in shared memory");
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, m->device);
  const int grid = (int)std::min<int64_t>(n_win, n_sm);
  if ((int64_t)grid * max_cells * 9 > (48ll << 30)) return fail(PUP_E_OOM, "pup_accumulate_rescaled: snippet scratch too large");
  const int32_t *d_r0 = r0, *d_c0 = c0, *d_h = h, *d_w = w, *d_slot = slot, *d_mode = mode;
  bool host_inputs = false;
  auto stage = [&](const int32_t*& d, const int32_t* src) -> cudaError_t {
    if (!src || is_device_ptr(src)) return cudaSuccess;
    host_inputs = true;
    int32_t* t;
    cudaError_t e = tmp.alloc((void**)&t, (size_t)n_win * 4);
    if (e != cudaSuccess) return e;
    d = t;
    return cudaMemcpyAsync(t, src, (size_t)n_win * 4, cudaMemcpyHostToDevice, st);
  };
  CK(stage(d_r0, r0));
  CK(stage(d_c0, c0));
  CK(stage(d_h, h));
  CK(stage(d_w, w));
  CK(stage(d_slot, slot));
  CK(stage(d_mode, mode));
  RescaleParams rp;
  rp.sp = StripeParams{m->pix, m->prow, m->bad, (m->flags & PUP_F_OOE) ? m->expected : nullptr, m->nb, rs,
                       m->ignore_diags, m->lr, m->flags};
  rp.expected = m->expected;
  rp.coverage = m->coverage;
  rp.r0 = d_r0;
  rp.c0 = d_c0;
  rp.h = d_h;
  rp.w = d_w;
  rp.slot = d_slot;
  rp.mode = d_mode;
  rp.n_win = n_win;
  rp.rs = rs;
  rp.n_slots = n_slots;
  rp.max_tmp = max_tmp;
  rp.flags = flags;
  rp.max_cells = max_cells;
  rp.acc = acc;
  CK(tmp.alloc((void**)&rp.scratch_d, (size_t)grid * max_cells * 8));
  CK(tmp.alloc((void**)&rp.scratch_m, (size_t)grid * max_cells));
  CK(tmp.alloc((void**)&rp.work, 16));
  CK(zero_async(rp.work, 16, st));
  CK(cudaFuncSetAttribute(k_rescale, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    SpanGuard span(2, st);  // timing tag of the main kernel
      k_rescale<<<grid, 1024, smem, st>>>(rp);
    LAUNCH_CHECK("k_rescale");
  }
  if (n_valid_out) {
    int32_t nv = 0;
    CK(cudaMemcpyAsync(&nv, rp.work + 1, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_valid_out = nv;
  } else if (host_inputs && !async) {
    CK(cudaStreamSynchronize(st));
  }
  return PUP_OK;
}

// ------------------------------------------------------------------------------------------ device-side windows
struct pup_rng {
  int device;
  cudaStream_t stream;  // the stream the state was allocated on
  pup_rng_state* st;
};

int pup_rng_create(int device, const uint32_t* key, int pos, void* stream, pup_rng_t** out) {
  if (!out || !key || pos < 0 || pos > MT_N) return fail(PUP_E_ARG, "pup_rng_create: bad arguments");
  *out = nullptr;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_rng_create: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  pup_rng_state h;
  memcpy(h.key, key, sizeof h.key);
  h.pos = pos;
  pup_rng* r = new pup_rng();
  r->device = device;
  r->stream = st;
  // stream-ordered allocation (no device-wide synchronisation like cudaMalloc / cudaFree); the 2.5 KB pageable source
  // on this frame is staged by the runtime before cudaMemcpyAsync returns
  cudaError_t e = cudaMallocAsync((void**)&r->st, sizeof(pup_rng_state), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(r->st, &h, sizeof h, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    if (r->st) cudaFreeAsync(r->st, st);
    delete r;
    return fail(PUP_E_CUDA, "pup_rng_create", e);
  }
  *out = r;
  return PUP_OK;
}

int pup_rng_read(pup_rng_t* r, uint32_t* key, int* pos, void* stream) {
  if (!r || !key || !pos) return fail(PUP_E_ARG, "pup_rng_read: bad arguments");
  DeviceGuard guard(r->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_rng_read: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  pup_rng_state h;
  CK(cudaMemcpyAsync(&h, r->st, sizeof h, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(key, h.key, sizeof h.key);
  *pos = h.pos;
  return PUP_OK;
}

int pup_rng_destroy(pup_rng_t* r) {
  if (!r) return PUP_OK;
  DeviceGuard guard(r->device);
  if (r->st) cudaFreeAsync(r->st, r->stream);
  delete r;
  return PUP_OK;
}

int pup_control_shifts(pup_rng_t* r, int64_t n_segments, const int64_t* seg_n, int64_t minshift, int64_t maxshift,
                       double resolution, int32_t* dbin, void* stream) {
  if (!r || n_segments < 0 || (n_segments > 0 && !seg_n)) return fail(PUP_E_ARG, "pup_control_shifts: bad arguments");
  const int64_t range = maxshift - 1 - minshift;
  if (range <= 0 || range >= 0xffffffffll || minshift <= -(1ll << 31) || maxshift >= (1ll << 31) || !(resolution > 0))
    return fail(PUP_E_ARG, "pup_control_shifts: need minshift + 1 < maxshift, both inside int32, and a positive resolution");
  if (dbin && !is_device_ptr(dbin)) return fail(PUP_E_ARG, "pup_control_shifts: dbin must be device memory");
  if (n_segments == 0) return PUP_OK;
  for (int64_t s = 0; s < n_segments; ++s)
    if (seg_n[s] <= 0) return fail(PUP_E_ARG, "pup_control_shifts: segment sizes must be positive");
  DeviceGuard guard(r->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_control_shifts: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch tmp(st);
  int64_t* d_seg;
  CK(tmp.alloc((void**)&d_seg, (size_t)n_segments * 8));
  // pageable source: the runtime stages it before returning, so seg_n may be freed by the caller right away
  CK(cudaMemcpyAsync(d_seg, seg_n, (size_t)n_segments * 8, cudaMemcpyHostToDevice, st));
  uint32_t mask = (uint32_t)range;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  k_mt_shifts<<<1, MT_THREADS, 0, st>>>(r->st, d_seg, n_segments, minshift, (uint32_t)range, mask, resolution, dbin);
  LAUNCH_CHECK("k_mt_shifts");
  return PUP_OK;
}

int pup_pair_windows_device(int device, int32_t m, const int32_t* stbin, const double* center, double mindist,
                            double maxdist, int32_t nctrl, const int64_t* per_offset, const int32_t* dbin, int32_t nb,
                            int W, const int64_t* key1, const int64_t* key2, const double* band_edges, int32_t n_edges,
                            int64_t band_weight, int flip_mode, int swap_on_flip, const int32_t* flipval,
                            const int32_t* ident, int nk, int nf, int32_t k_lo, int32_t k_hi,
                            const int64_t* per_offset_part, int32_t region_index, int32_t* r0, int32_t* c0,
                            int32_t* slot, uint64_t* first_seen, uint64_t* n_roi, void* stream) {
  if (m < 0 || nctrl < 0 || W <= 0 || nb <= 0 || nk <= 0 || nf <= 0 || k_lo < 0 || k_hi < k_lo || k_hi > m ||
      region_index < 0 || region_index >= (1 << 20))
    return fail(PUP_E_ARG, "pup_pair_windows_device: bad sizes");
  if (m < 2) return PUP_OK;
  if (!stbin || !center || !per_offset || !r0 || !c0 || !slot)
    return fail(PUP_E_ARG, "pup_pair_windows_device: null arrays");
  if (flip_mode < 0 || flip_mode > 2 || (flip_mode != 0 && !flipval) || (band_edges && n_edges <= 0))
    return fail(PUP_E_ARG, "pup_pair_windows_device: bad flip / band arguments");
  if (!is_device_ptr(r0) || !is_device_ptr(c0) || !is_device_ptr(slot) || (dbin && !is_device_ptr(dbin)) ||
      (first_seen && !is_device_ptr(first_seen)) || (n_roi && !is_device_ptr(n_roi)))
    return fail(PUP_E_ARG, "pup_pair_windows_device: outputs and dbin must be device memory");
  if (!per_offset_part) per_offset_part = per_offset;  // the whole region
  int64_t total = 0, total_part = 0;
  std::vector<int64_t> base((size_t)m), base_part((size_t)m);
  for (int32_t i = 0; i < m; ++i) {
    base[(size_t)i] = total;
    base_part[(size_t)i] = total_part;
    total += per_offset[i];
    total_part += per_offset_part[i];
    if (per_offset_part[i] > per_offset[i]) return fail(PUP_E_ARG, "pup_pair_windows_device: part counts exceed the region's");
  }
  if (total == 0 || total_part == 0) return PUP_OK;
  if (nctrl > 0 && !dbin) return fail(PUP_E_ARG, "pup_pair_windows_device: control shifts missing");
  if (total * (1 + (int64_t)nctrl) >= (1ll << 40)) return fail(PUP_E_ARG, "pup_pair_windows_device: too many windows");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_pair_windows_device: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch tmp(st);
  // small per-feature inputs: staged on the caller's stream (pageable sources are consumed before the call returns)
  auto stage = [&](const void* src, size_t bytes, const void** dst) -> int {
    *dst = src;
    if (!src || bytes == 0 || is_device_ptr(src)) return PUP_OK;
    void* t;
    CK(tmp.alloc(&t, bytes));
    CK(cudaMemcpyAsync(t, src, bytes, cudaMemcpyHostToDevice, st));
    *dst = t;
    return PUP_OK;
  };
  const void *d_st, *d_ce, *d_po, *d_base, *d_pop, *d_basep, *d_k1, *d_k2, *d_ed, *d_fv, *d_id;
  int rc;
  if ((rc = stage(stbin, (size_t)m * 4, &d_st)) != PUP_OK) return rc;
  if ((rc = stage(center, (size_t)m * 8, &d_ce)) != PUP_OK) return rc;
  if ((rc = stage(per_offset, (size_t)m * 8, &d_po)) != PUP_OK) return rc;
  if ((rc = stage(base.data(), (size_t)m * 8, &d_base)) != PUP_OK) return rc;
  if ((rc = stage(per_offset_part, (size_t)m * 8, &d_pop)) != PUP_OK) return rc;
  if ((rc = stage(base_part.data(), (size_t)m * 8, &d_basep)) != PUP_OK) return rc;
  if ((rc = stage(key1, (size_t)m * 8, &d_k1)) != PUP_OK) return rc;
  if ((rc = stage(key2, (size_t)m * 8, &d_k2)) != PUP_OK) return rc;
  if ((rc = stage(band_edges, (size_t)n_edges * 8, &d_ed)) != PUP_OK) return rc;
  if ((rc = stage(flipval, (size_t)m * 4, &d_fv)) != PUP_OK) return rc;
  if ((rc = stage(ident, (size_t)m * 4, &d_id)) != PUP_OK) return rc;
  PairGen g;
  g.m = m;
  g.nctrl = nctrl;
  g.W = W;
  g.nb = nb;
  g.nk = nk;
  g.nf = nf;
  g.targets = ident ? 2 : 1;
  g.k_lo = k_lo;
  g.k_hi = k_hi;
  g.region_index = region_index;
  g.stbin = (const int32_t*)d_st;
  g.center = (const double*)d_ce;
  g.mindist = mindist;
  g.maxdist = maxdist;
  g.base = (const int64_t*)d_base;
  g.per_offset = (const int64_t*)d_po;
  g.base_part = (const int64_t*)d_basep;
  g.per_offset_part = (const int64_t*)d_pop;
  g.dbin = dbin;
  g.key1 = (const int64_t*)d_k1;
  g.key2 = (const int64_t*)d_k2;
  g.edges = (const double*)d_ed;
  g.n_edges = band_edges ? n_edges : 0;
  g.band_weight = band_weight;
  g.flip_mode = flip_mode;
  g.swap_on_flip = swap_on_flip;
  g.flipval = (const int32_t*)d_fv;
  g.ident = (const int32_t*)d_id;
  g.r0 = r0;
  g.c0 = c0;
  g.slot = slot;
  g.first = (unsigned long long*)first_seen;
  g.n_roi = (unsigned long long*)n_roi;
  k_pair_windows<<<m - 1, 256, 0, st>>>(g);
  LAUNCH_CHECK("k_pair_windows");
  // no synchronisation: the pageable host sources (incl. `base` on this frame) have been copied to the driver's
  // staging memory when cudaMemcpyAsync returns
  return PUP_OK;
}

int pup_algorithmic_bytes(const pup_region_t* m, int64_t n_win, const int32_t* r0, const int32_t* c0, int W,
                          unsigned flags, void* stream, int64_t* bytes_out, int64_t* nnz_out) {
  if (!m || n_win < 0 || W <= 0 || !bytes_out) return fail(PUP_E_ARG, "pup_algorithmic_bytes: bad arguments");
  DeviceGuard guard(m->device);
  if (!guard.ok) return fail(PUP_E_NODEV, "pup_algorithmic_bytes: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch tmp(st);
  const int32_t *d_r0 = r0, *d_c0 = c0;
  if (n_win > 0 && !is_device_ptr(r0)) {
    int32_t* t;
    CK(tmp.alloc((void**)&t, (size_t)n_win * 4));
    CK(cudaMemcpyAsync(t, r0, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
    d_r0 = t;
  }
  if (n_win > 0 && !is_device_ptr(c0)) {
    int32_t* t;
    CK(tmp.alloc((void**)&t, (size_t)n_win * 4));
    CK(cudaMemcpyAsync(t, c0, (size_t)n_win * 4, cudaMemcpyHostToDevice, st));
    d_c0 = t;
  }
  unsigned long long* d_out;
  CK(tmp.alloc((void**)&d_out, 16));
  CK(zero_async(d_out, 16, st));
  if (n_win > 0) {
    k_count_nnz<<<148 * 8, 256, 0, st>>>(m->pix, m->prow, d_r0, d_c0, n_win, m->nb, W, m->lr, d_out, d_out + 1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PUP_E_CUDA, "launch k_count_nnz", e);
  }
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, d_out, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  int64_t per_win = 16 + (int64_t)(W + 1) * 4;
  if (m->bad) per_win += 16ll * W;
  if (flags & PUP_F_COVERAGE) per_win += 16ll * W;
  *bytes_out = (int64_t)h[1] * per_win + (int64_t)h[0] * 8;
  if (nnz_out) *nnz_out = (int64_t)h[0];
  return PUP_OK;
}

}  // extern "C"
