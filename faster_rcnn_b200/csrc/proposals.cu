// K-a: anchors + delta decode + sanitize + validity filter + radix-select top-k + sort.
//
// Kernel 1 (decode_kernel): one thread per anchor, fully parallel over (anchor, image).
//   Reads regr (16 B) + cls (4 B), writes one 64-bit sort key and one packed int16x4 box
//   (16 B) per anchor.  HBM-bound, coalesced 128-bit loads/stores.
// Kernel 2 (topk_kernel): one CTA per image.  MSB-first 11-bit radix select over the
//   64-bit keys (unique because the anchor index is part of the key) finds the exact
//   k-th largest key, survivors are compacted into shared memory, sorted descending with a
//   register-blocked bitonic network and gathered into the output arrays.
//
// Reference semantics (file:line under /root/reference/faster_rcnn):
//   det_util.py:162-175 anchors (centre = cell index, x1 = x - w//2, x2 = x1 + w)
//   util.py:111-142     float32 decode, separate roundings (this TU is built with -fmad=false)
//   det_util.py:179-192 sanitize order, :196-205 validity, :68-76/:147-155 sort + top-k + int16
#include "common.cuh"

namespace frcnn {

struct __align__(8) BoxI16 { short x1, y1, x2, y2; };

constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_WARPS = TOPK_THREADS / 32;
constexpr int TOPK_BITS = 11;                       // radix-select digit width
constexpr int TOPK_BINS = 1 << TOPK_BITS;
constexpr int TOPK_META = TOPK_BINS + 1;            // per image: [valid count, histogram of the keys' top 11 bits]

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ regr,
                                                     const float* __restrict__ cls, AnchorTable tab,
                                                     int rows, int cols, int n_per_image,
                                                     unsigned long long* __restrict__ keys,
                                                     BoxI16* __restrict__ boxes,
                                                     float4* __restrict__ dense,
                                                     int* __restrict__ valid_count) {
  const int img = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = false;
  unsigned digit = 0xffffffffu;
  if (i < n_per_image) {
    const size_t g = (size_t)img * n_per_image + i;
    const int a = i % tab.n;
    const int loc = i / tab.n;
    const int cx_i = loc % cols, cy_i = loc / cols;
    const int aw = tab.w[a], ah = tab.h[a];

    // anchors are integer valued -> exact in float32
    float x = (float)(cx_i - (aw >> 1));
    float y = (float)(cy_i - (ah >> 1));
    float w = (float)aw;     // (x + aw) - x
    float hgt = (float)ah;

    const float4 r = ldg_f4(regr + 4 * g);
    const float tx = __fdiv_rn(r.x, 10.0f), ty = __fdiv_rn(r.y, 10.0f);
    const float tw = __fdiv_rn(r.z, 5.0f), th = __fdiv_rn(r.w, 5.0f);

    x = __fadd_rn(x, __fdiv_rn(w, 2.0f));
    y = __fadd_rn(y, __fdiv_rn(hgt, 2.0f));
    x = __fadd_rn(x, __fmul_rn(tx, w));
    y = __fadd_rn(y, __fmul_rn(ty, hgt));
    w = __fmul_rn(w, np_expf(tw));
    hgt = __fmul_rn(hgt, np_expf(th));
    x = __fsub_rn(x, __fdiv_rn(w, 2.0f));
    y = __fsub_rn(y, __fdiv_rn(hgt, 2.0f));
    x = rintf(x); y = rintf(y); w = rintf(w); hgt = rintf(hgt);     // np.round: half to even
    float x2 = __fadd_rn(w, x), y2 = __fadd_rn(hgt, y);

    x2 = np_max(__fadd_rn(x, 1.0f), x2);
    y2 = np_max(__fadd_rn(y, 1.0f), y2);
    x = np_max(0.0f, x);
    y = np_max(0.0f, y);
    x2 = np_min((float)(cols - 1), x2);
    y2 = np_min((float)(rows - 1), y2);

    if (dense) dense[g] = make_float4(x, y, x2, y2);

    valid = (x2 > x) && (y2 > y);
    unsigned long long key = 0ull;
    BoxI16 b = {0, 0, 0, 0};
    if (valid) {
      key = ((unsigned long long)mono_key(__ldg(cls + g)) << 32) | (unsigned)i;
      b.x1 = (short)(int)x; b.y1 = (short)(int)y; b.x2 = (short)(int)x2; b.y2 = (short)(int)y2;
    }
    keys[g] = key;
    boxes[g] = b;
    digit = valid ? (unsigned)(key >> (64 - TOPK_BITS)) : 0xffffffffu;
  }
  // first pass of top-k's radix select, done here while the key is in a register: histogram of the top 11 bits
  // (a handful of distinct digits per warp -> one aggregated atomic per digit)
  int* meta = valid_count + (size_t)img * TOPK_META;
  const unsigned peers = __match_any_sync(0xffffffffu, digit);
  if (valid && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(meta + 1 + digit, __popc(peers));
  // number of valid anchors of the image (top-k needs it): one atomic per CTA
  const int nv = __syncthreads_count(valid);
  if (threadIdx.x == 0 && nv) atomicAdd(meta, nv);
}


// block-wide inclusive scan of one int per thread (1024 threads); `warp_tot` is [TOPK_WARPS] shared
__device__ __forceinline__ int block_inclusive_scan(int v, int* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += u;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += u;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  const int before = warp ? warp_tot[warp - 1] : 0;
  __syncthreads();                                   // warp_tot is reused by the next call
  return v + before;
}

// shared-memory slot of logical element e: one pad entry per E elements keeps the blocked
// register <-> shared transfers (thread t owns e = t*E .. t*E+E-1) at the 2-way minimum of 64-bit accesses
template <int E>
__device__ __forceinline__ int pad_slot(int e) { return e + e / E; }

// Block-wide exact selection of up to TWO ranks in one set of sweeps: splitter T_s such that exactly `rank_s` keys are
// >= T_s (keys are unique and non-zero; the caller guarantees 0 < rank < number of valid keys).  MSB-first radix
// select, 11-bit digits, shared histograms fed by warp-aggregated atomics (__match_any_sync: one atomic per distinct
// digit per warp), suffix scan over the bins.  The pass at which a splitter's bucket holds exactly the keys still
// needed ends that select (3 passes for tie-free float scores).  A CTA in the middle of the rank range needs both of
// its slice bounds; selecting them jointly reads the keys once per pass instead of twice, and while the two prefixes
// still agree (always in the first pass) one histogram serves both.
// Keys are fetched 8 per thread before any is consumed, which hides the L2 latency.
constexpr int TOPK_U = 8;
__device__ void radix_select2(const unsigned long long* __restrict__ keys, const int* __restrict__ g_hist1, int n,
                              int rank_a, int rank_b, bool use_b,
                              unsigned* hist /* [2][TOPK_BINS] */, int* warp_tot, unsigned long long* s_prefix /* [2] */,
                              int* s_need /* [2] */, int* s_done /* [2] */, unsigned long long& t_a,
                              unsigned long long& t_b) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_prefix[0] = s_prefix[1] = 0ull;
    s_need[0] = rank_a; s_need[1] = rank_b;
    s_done[0] = 0; s_done[1] = use_b ? 0 : 1;
  }
  __syncthreads();
  for (int hi = 64; hi > 0;) {
    const int bits = hi < TOPK_BITS ? hi : TOPK_BITS;
    const int shift = hi - bits;
    for (int i = tid; i < 2 * TOPK_BINS; i += TOPK_THREADS) hist[i] = 0u;
    __syncthreads();
    const unsigned long long pa = s_prefix[0], pb = s_prefix[1];
    const bool da = s_done[0] != 0, db = s_done[1] != 0;
    const bool same = !da && !db && (hi == 64 || (pa >> hi) == (pb >> hi));   // identical histograms: build A's only
    if (hi == 64) {                                  // first pass: decode_kernel already histogrammed the top bits
      for (int i = tid; i < TOPK_BINS; i += TOPK_THREADS) hist[i] = (unsigned)__ldg(g_hist1 + i);
    } else
    for (int base = 0; base < n; base += TOPK_THREADS * TOPK_U) {
      unsigned long long kk[TOPK_U];
#pragma unroll
      for (int u = 0; u < TOPK_U; ++u) {
        const int i = base + u * TOPK_THREADS + tid;
        kk[u] = (i < n) ? __ldg(keys + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < TOPK_U; ++u) {
        const unsigned long long key = kk[u];
        const bool in_a = !da && key != 0ull && (hi == 64 || (key >> hi) == (pa >> hi));
        const unsigned digit = (unsigned)((key >> shift) & (unsigned long long)((1u << bits) - 1u));
        if (in_a) atomicAdd(&hist[digit], 1u);       // few keys are left after the first pass and their digits are spread out
        if (!db && !same && key != 0ull && (key >> hi) == (pb >> hi)) atomicAdd(&hist[TOPK_BINS + digit], 1u);
      }
    }
    __syncthreads();
#pragma unroll
    for (int sel = 0; sel < 2; ++sel) {
      if (sel == 0 ? da : db) continue;              // block-uniform
      const unsigned* hh = hist + ((sel == 1 && !same) ? TOPK_BINS : 0);
      const unsigned long long prefix = sel == 0 ? pa : pb;
      const int need = s_need[sel];
      // suffix scan from the top bin: thread t owns bins BINS-1-2t (c0) and BINS-2-2t (c1)
      const int c0 = (int)hh[TOPK_BINS - 1 - 2 * tid], c1 = (int)hh[TOPK_BINS - 2 - 2 * tid];
      const int incl = block_inclusive_scan(c0 + c1, warp_tot);
      const int excl = incl - c0 - c1;
      int d = -1, rest = 0, cnt = 0;
      if (excl < need && excl + c0 >= need) { d = TOPK_BINS - 1 - 2 * tid; rest = need - excl; cnt = c0; }
      else if (excl + c0 < need && incl >= need) { d = TOPK_BINS - 2 - 2 * tid; rest = need - excl - c0; cnt = c1; }
      if (d >= 0) {                                  // exactly one thread
        s_prefix[sel] = prefix | ((unsigned long long)d << shift);
        s_need[sel] = rest;
        if (shift == 0 || cnt == rest) s_done[sel] = 1;   // bucket taken whole: the low bits are free
      }
    }
    __syncthreads();
    hi = shift;
    if (s_done[0] && s_done[1]) break;
  }
  t_a = s_prefix[0];
  t_b = s_prefix[1];
  __syncthreads();
}

// Suffix scan over a TOPK_BINS histogram: the bin that holds the need-th largest element (1-based).
// out[0] = bin, out[1] = rank of that element inside the bin, out[2] = size of the bin.  Block-wide; ends with a barrier.
__device__ __forceinline__ void find_bucket(const unsigned* hh, int need, int* warp_tot, int* out) {
  const int tid = threadIdx.x;
  const int c0 = (int)hh[TOPK_BINS - 1 - 2 * tid], c1 = (int)hh[TOPK_BINS - 2 - 2 * tid];
  const int incl = block_inclusive_scan(c0 + c1, warp_tot);
  const int excl = incl - c0 - c1;
  if (excl < need && excl + c0 >= need) { out[0] = TOPK_BINS - 1 - 2 * tid; out[1] = need - excl; out[2] = c0; }
  else if (excl + c0 < need && incl >= need) { out[0] = TOPK_BINS - 2 - 2 * tid; out[1] = need - excl - c0; out[2] = c1; }
  __syncthreads();
}

// Exact selection inside ONE top-digit bucket whose keys sit in shared memory (`st`, n_st keys, all with the same top
// TOPK_BITS bits = prefix >> 53): the key T with exactly `need` bucket keys >= T.  11-bit passes over the stash only --
// no global sweep; one or two passes for float scores.
__device__ unsigned long long select_in_stash(const unsigned long long* st, int n_st, unsigned long long prefix, int need,
                                              unsigned* hist, int* warp_tot, int* s_out) {
  int hi = 64 - TOPK_BITS;
  while (true) {
    const int bits = hi < TOPK_BITS ? hi : TOPK_BITS, shift = hi - bits;
    for (int i = threadIdx.x; i < TOPK_BINS; i += TOPK_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n_st; i += TOPK_THREADS) {
      const unsigned long long key = st[i];
      if ((key >> hi) == (prefix >> hi)) atomicAdd(&hist[(unsigned)((key >> shift) & (unsigned long long)((1u << bits) - 1u))], 1u);
    }
    __syncthreads();
    find_bucket(hist, need, warp_tot, s_out);
    const int d = s_out[0], rest = s_out[1], cnt = s_out[2];
    prefix |= (unsigned long long)d << shift;
    need = rest;
    hi = shift;
    if (shift == 0 || cnt == rest) break;            // bucket taken whole: the low bits are free
  }
  __syncthreads();
  return prefix;
}

constexpr int TOPK_STASH = 6144;                     // boundary-bucket keys a CTA can resolve in shared memory (48 KB)

// `splits` CTAs per image.  CTA r produces ranks [r*k/splits, (r+1)*k/splits) of the descending top-k
// order on its own: it selects the two splitters that bound its slice (exact radix select over all
// keys of the image, L2-resident), compacts the keys between them into shared memory and sorts them
// with a bitonic network that keeps E keys per thread in registers: strides < E are register swaps,
// strides < 32E warp shuffles, only strides >= 32E go through shared memory.  The slices need no
// merge and no inter-CTA communication, and the n log^2 n sort work shrinks with the slice.
template <int E>
__global__ void __launch_bounds__(TOPK_THREADS, 1)
topk_kernel(const unsigned long long* __restrict__ keys_all, const BoxI16* __restrict__ boxes_all,
            const float* __restrict__ cls_all, const int* __restrict__ valid_count, int n, int k, int splits,
            BoxI16* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_index,
            int* __restrict__ out_count) {
  constexpr int M = TOPK_THREADS * E;                 // sort size (power of two), >= slice length
  extern __shared__ __align__(16) unsigned long long sbuf[];      // pad_slot<E>(M) entries
  __shared__ unsigned hist[2 * TOPK_BINS];
  __shared__ int warp_tot[TOPK_WARPS];
  __shared__ unsigned long long s_prefix[2];
  __shared__ int s_need[2], s_done[2], s_count, s_sel[4], s_stash[2];

  const int img = blockIdx.x / splits, part = blockIdx.x - img * splits;
  const unsigned long long* keys = keys_all + (size_t)img * n;
  const int tid = threadIdx.x, lane = tid & 31;

  if (tid == 0) s_count = 0;
  __syncthreads();
  const int* meta = valid_count + (size_t)img * TOPK_META;
  const int n_valid = __ldg(meta);                    // counted by decode_kernel
  const int m = min(k, n_valid);
  const int first = (int)((long long)part * k / splits);                       // ranks [first, last) belong to this CTA
  const int slice_end = (int)((long long)(part + 1) * k / splits);
  const int last = min(slice_end, m);

  unsigned long long t_hi = ~0ull, t_lo = 1ull;
  // Slice membership.  decode_kernel histogrammed the keys' top 11 bits; the buckets of the two slice bounds follow from
  // that histogram alone.  ONE sweep over the image's keys then sends every key strictly between the two boundary
  // buckets straight into the slice and parks the keys OF the boundary buckets in a shared-memory stash; the exact
  // splitters are resolved inside the stash (select_in_stash) and the stash keys on the right side of them join the slice.
  // The first version swept the keys three times (two select passes + compaction), each sweep an L2 round trip chain;
  // it remains as the fall-back when the boundary buckets outgrow the stash (scores crowded into one quarter-octave).
  unsigned long long* stash = sbuf + (M + TOPK_THREADS);
  bool done = first >= last;
  if (!done) {
    const bool need_hi = first > 0, need_lo = last < n_valid;
    for (int i = tid; i < TOPK_BINS; i += TOPK_THREADS) hist[i] = (unsigned)__ldg(meta + 1 + i);
    __syncthreads();
    int da = TOPK_BINS, rest_a = 0, cnt_a = 0, db = -1, rest_b = 0, cnt_b = 0;
    if (need_hi) {
      find_bucket(hist, first, warp_tot, s_sel);
      da = s_sel[0]; rest_a = s_sel[1]; cnt_a = s_sel[2];
      __syncthreads();
    }
    if (need_lo) {
      find_bucket(hist, last, warp_tot, s_sel);
      db = s_sel[0]; rest_b = s_sel[1]; cnt_b = s_sel[2];
      __syncthreads();
    }
    const bool one_bucket = da == db;
    const int n_stash_a = cnt_a, n_stash_b = one_bucket ? 0 : cnt_b;
    if (n_stash_a + n_stash_b <= TOPK_STASH) {
      done = true;
      if (tid == 0) { s_stash[0] = 0; s_stash[1] = 0; }
      __syncthreads();
      for (int base = 0; base < n; base += TOPK_THREADS * TOPK_U) {
        unsigned long long kk[TOPK_U];
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
          const int i = base + u * TOPK_THREADS + tid;
          kk[u] = (i < n) ? __ldg(keys + i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < TOPK_U; ++u) {
          const int dg = (int)(kk[u] >> (64 - TOPK_BITS));
          const bool valid = kk[u] != 0ull;
          const bool take = valid && dg < da && dg > db;
          const unsigned ballot = __ballot_sync(0xffffffffu, take);
          int wbase = 0;
          if (lane == 0 && ballot) wbase = atomicAdd(&s_count, __popc(ballot));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
          if (take) sbuf[pad_slot<E>(wbase + __popc(ballot & ((1u << lane) - 1u)))] = kk[u];
          if (valid && dg == da) stash[atomicAdd(&s_stash[0], 1)] = kk[u];
          else if (valid && dg == db) stash[n_stash_a + atomicAdd(&s_stash[1], 1)] = kk[u];
        }
      }
      __syncthreads();
      if (need_hi) t_hi = select_in_stash(stash, n_stash_a, (unsigned long long)da << (64 - TOPK_BITS), rest_a, hist, warp_tot, s_sel);
      if (need_lo) t_lo = one_bucket ? select_in_stash(stash, n_stash_a, (unsigned long long)db << (64 - TOPK_BITS), rest_b, hist, warp_tot, s_sel)
                                     : select_in_stash(stash + n_stash_a, n_stash_b, (unsigned long long)db << (64 - TOPK_BITS), rest_b, hist, warp_tot, s_sel);
      for (int i = tid; i < n_stash_a + n_stash_b; i += TOPK_THREADS) {
        const unsigned long long key = stash[i];
        if (key >= t_lo && (first == 0 || key < t_hi)) sbuf[pad_slot<E>(atomicAdd(&s_count, 1))] = key;
      }
    }
  }
  if (!done) {
    // fall-back: exact splitters by radix select over all keys (keys >= t_lo and < t_hi are exactly the ranks [first, last))
    const bool need_hi = first > 0, need_lo = last < n_valid;
    unsigned long long ta = 0ull, tb = 0ull;
    if (need_hi && need_lo) {
      radix_select2(keys, meta + 1, n, first, last, true, hist, warp_tot, s_prefix, s_need, s_done, ta, tb);
      t_hi = ta;
      t_lo = tb;
    } else if (need_hi) {
      radix_select2(keys, meta + 1, n, first, 0, false, hist, warp_tot, s_prefix, s_need, s_done, ta, tb);
      t_hi = ta;
    } else if (need_lo) {
      radix_select2(keys, meta + 1, n, last, 0, false, hist, warp_tot, s_prefix, s_need, s_done, ta, tb);
      t_lo = ta;
    }
    // compaction of the slice into shared memory (order irrelevant, sorted next)
    for (int base = 0; base < n; base += TOPK_THREADS * TOPK_U) {
      unsigned long long kk[TOPK_U];
#pragma unroll
      for (int u = 0; u < TOPK_U; ++u) {
        const int i = base + u * TOPK_THREADS + tid;
        kk[u] = (i < n) ? __ldg(keys + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < TOPK_U; ++u) {
        const bool take = kk[u] != 0ull && kk[u] >= t_lo && (first == 0 || kk[u] < t_hi);
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        int wbase = 0;
        if (lane == 0 && ballot) wbase = atomicAdd(&s_count, __popc(ballot));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (take) sbuf[pad_slot<E>(wbase + __popc(ballot & ((1u << lane) - 1u)))] = kk[u];
      }
    }
  }
  __syncthreads();
  for (int i = s_count + tid; i < M; i += TOPK_THREADS) sbuf[pad_slot<E>(i)] = 0ull;   // pad with the smallest key
  __syncthreads();

  // ---- bitonic sort, descending, E keys per thread in registers ----
  unsigned long long v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) v[i] = sbuf[pad_slot<E>(tid * E + i)];
  for (int size = 2; size <= M; size <<= 1) {
    int stride = size >> 1;
    if (stride >= 32 * E) {
      // cross-warp strides of this merge level: classic shared-memory passes
      __syncthreads();
#pragma unroll
      for (int i = 0; i < E; ++i) sbuf[pad_slot<E>(tid * E + i)] = v[i];
      for (; stride >= 32 * E; stride >>= 1) {
        __syncthreads();
        for (int t = tid; t < (M >> 1); t += TOPK_THREADS) {
          const int lo = 2 * t - (t & (stride - 1)), hi2 = lo + stride;
          const bool desc = ((lo & size) == 0);
          const unsigned long long a = sbuf[pad_slot<E>(lo)], b = sbuf[pad_slot<E>(hi2)];
          if ((a < b) == desc) { sbuf[pad_slot<E>(lo)] = b; sbuf[pad_slot<E>(hi2)] = a; }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < E; ++i) v[i] = sbuf[pad_slot<E>(tid * E + i)];
    }
    for (; stride >= E; stride >>= 1) {              // partner in another lane of the warp
      const int lane_xor = stride / E;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[i], lane_xor);
        const int e = tid * E + i;
        const bool want_max = ((e & stride) == 0) == ((e & size) == 0);
        v[i] = want_max ? (v[i] > other ? v[i] : other) : (v[i] < other ? v[i] : other);
      }
    }
#pragma unroll
    for (int st = E >> 1; st > 0; st >>= 1) {        // partner in the same thread
      if (st <= stride) {
#pragma unroll
        for (int i = 0; i < E; ++i) {
          if ((i & st) == 0) {
            const int e = tid * E + i;
            const bool desc = ((e & size) == 0);
            const unsigned long long a = v[i], b = v[i | st];
            if ((a < b) == desc) { v[i] = b; v[i | st] = a; }
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < E; ++i) sbuf[pad_slot<E>(tid * E + i)] = v[i];
  __syncthreads();

  const BoxI16* boxes = boxes_all + (size_t)img * n;
  const float* cls = cls_all + (size_t)img * n;
  for (int r = first + tid; r < slice_end; r += TOPK_THREADS) {
    const size_t o = (size_t)img * k + r;
    if (r < last) {
      const int idx = (int)(unsigned)(sbuf[pad_slot<E>(r - first)] & 0xffffffffull);
      out_boxes[o] = boxes[idx];
      out_scores[o] = __ldg(cls + idx);
      out_index[o] = idx;
    } else {
      out_boxes[o] = BoxI16{0, 0, 0, 0};
      out_scores[o] = 0.0f;
      out_index[o] = -1;
    }
  }
  if (tid == 0 && part == 0) out_count[img] = m;
}

template <int E>
static int launch_topk(frcnn_handle* h, cudaStream_t stream, const unsigned long long* keys, const BoxI16* boxes,
                       const float* cls, const int* valid_count, int n, int k, int splits, int batch, BoxI16* out_boxes,
                       float* out_scores, int32_t* out_index, int32_t* out_count) {
  const size_t smem = (size_t)(TOPK_THREADS * E + TOPK_THREADS + TOPK_STASH) * sizeof(unsigned long long);
  if (smem + 12 * 1024 > (size_t)h->max_smem_optin)
    return fail(h, FRCNN_ERR_UNSUPPORTED, "decode_topk: k too large for the shared-memory sort%s%s");
  FRCNN_CUDA(h, cudaFuncSetAttribute(topk_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_kernel<E><<<batch * splits, TOPK_THREADS, smem, stream>>>(keys, boxes, cls, valid_count, n, k, splits, out_boxes,
                                                               out_scores, out_index, out_count);
  FRCNN_LAUNCH_CHECK(h, "topk_kernel");
  return FRCNN_OK;
}

int launch_decode_topk(frcnn_handle* h, cudaStream_t stream, const float* regr, const float* cls,
                       const AnchorTable& tab, int rows, int cols, int k, int batch,
                       int16_t* out_boxes, float* out_scores, int32_t* out_index,
                       int32_t* out_count, float* dense_boxes) {
  const int n = rows * cols * tab.n;
  if (k > TOPK_THREADS * 16)
    return fail(h, FRCNN_ERR_UNSUPPORTED, "decode_topk: k above 16384 is not supported%s%s");
  const size_t key_bytes = align_up((size_t)batch * n * sizeof(unsigned long long), 256);
  const size_t box_bytes = align_up((size_t)batch * n * sizeof(BoxI16), 256);
  void* ws = nullptr;
  int rc = arena_get(h, stream, key_bytes + box_bytes + (size_t)batch * TOPK_META * sizeof(int), &ws);
  if (rc) return rc;
  auto* keys = reinterpret_cast<unsigned long long*>(ws);
  auto* boxes = reinterpret_cast<BoxI16*>(reinterpret_cast<char*>(ws) + key_bytes);
  int* valid_count = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + key_bytes + box_bytes);
  FRCNN_CUDA(h, cudaMemsetAsync(valid_count, 0, (size_t)batch * TOPK_META * sizeof(int), stream));

  dim3 grid((n + 255) / 256, batch);
  decode_kernel<<<grid, 256, 0, stream>>>(regr, cls, tab, rows, cols, n, keys, boxes,
                                         reinterpret_cast<float4*>(dense_boxes), valid_count);
  FRCNN_LAUNCH_CHECK(h, "decode_kernel");
  // CTAs per image: as many rank slices as keep the GPU filled in one wave, slices of >= 896 keys (k = 8000 at batch 1:
  // eight 1000-key slices sorted one key per thread; same-box A/B 38 -> 34 us on clustered scores, equal on uniform ones)
  int splits = 1;
  while (splits < 8 && (long long)batch * splits * 2 <= h->sm_count && k / (splits * 2) >= 896) splits <<= 1;
  const int slice = (k + splits - 1) / splits + 1;
  auto* ob = reinterpret_cast<BoxI16*>(out_boxes);
  if (slice <= TOPK_THREADS * 1) return launch_topk<1>(h, stream, keys, boxes, cls, valid_count, n, k, splits, batch, ob, out_scores, out_index, out_count);
  if (slice <= TOPK_THREADS * 2) return launch_topk<2>(h, stream, keys, boxes, cls, valid_count, n, k, splits, batch, ob, out_scores, out_index, out_count);
  if (slice <= TOPK_THREADS * 4) return launch_topk<4>(h, stream, keys, boxes, cls, valid_count, n, k, splits, batch, ob, out_scores, out_index, out_count);
  if (slice <= TOPK_THREADS * 8) return launch_topk<8>(h, stream, keys, boxes, cls, valid_count, n, k, splits, batch, ob, out_scores, out_index, out_count);
  return launch_topk<16>(h, stream, keys, boxes, cls, valid_count, n, k, splits, batch, ob, out_scores, out_index, out_count);
}

}  // namespace frcnn
