// K-a: anchors + delta decode + sanitize + validity filter + radix-select top-k + sort.
//
// Kernel 1 (decode_kernel): one thread per anchor, fully parallel over (anchor, image).
//   Reads regr (16 B) + cls (4 B), writes one 64-bit sort key and one packed int16x4 box
//   (16 B) per anchor.  HBM-bound, coalesced 128-bit loads/stores.
// Kernel 2 (topk_kernel): one CTA per image.  MSB-first 8-bit radix select over the
//   64-bit keys (unique because the anchor index is part of the key) finds the exact
//   k-th largest key, survivors are compacted into shared memory, bitonic-sorted
//   descending and gathered into the output arrays.
//
// Reference semantics (file:line under /root/reference/faster_rcnn):
//   det_util.py:162-175 anchors (centre = cell index, x1 = x - w//2, x2 = x1 + w)
//   util.py:111-142     float32 decode, separate roundings (this TU is built with -fmad=false)
//   det_util.py:179-192 sanitize order, :196-205 validity, :68-76/:147-155 sort + top-k + int16
#include "common.cuh"

namespace frcnn {

struct __align__(8) BoxI16 { short x1, y1, x2, y2; };

__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ regr,
                                                     const float* __restrict__ cls, AnchorTable tab,
                                                     int rows, int cols, int n_per_image,
                                                     unsigned long long* __restrict__ keys,
                                                     BoxI16* __restrict__ boxes,
                                                     float4* __restrict__ dense) {
  const int img = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_per_image) return;
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
  w = __fmul_rn(w, expf(tw));
  hgt = __fmul_rn(hgt, expf(th));
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

  const bool valid = (x2 > x) && (y2 > y);
  unsigned long long key = 0ull;
  BoxI16 b = {0, 0, 0, 0};
  if (valid) {
    key = ((unsigned long long)mono_key(__ldg(cls + g)) << 32) | (unsigned)i;
    b.x1 = (short)(int)x; b.y1 = (short)(int)y; b.x2 = (short)(int)x2; b.y2 = (short)(int)y2;
  }
  keys[g] = key;
  boxes[g] = b;
}

constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_WARPS = TOPK_THREADS / 32;

// One CTA per image.  Dynamic smem: m_pow2 * 8 bytes of sort buffer.
__global__ void __launch_bounds__(TOPK_THREADS, 1)
topk_kernel(const unsigned long long* __restrict__ keys_all, const BoxI16* __restrict__ boxes_all,
            const float* __restrict__ cls_all, int n, int k, int m_pow2,
            BoxI16* __restrict__ out_boxes, float* __restrict__ out_scores,
            int* __restrict__ out_index, int* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned long long sbuf[];
  __shared__ unsigned hist[TOPK_WARPS][256];
  __shared__ unsigned total[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_count, s_valid, s_done;

  const int img = blockIdx.x;
  const unsigned long long* keys = keys_all + (size_t)img * n;
  const int tid = threadIdx.x, warp = tid >> 5;

  // pass 0: number of valid anchors
  int local_valid = 0;
  for (int i = tid; i < n; i += TOPK_THREADS) local_valid += (keys[i] != 0ull);
  local_valid = __reduce_add_sync(0xffffffffu, local_valid);
  if (tid == 0) { s_valid = 0; s_count = 0; s_done = 0; }
  __syncthreads();
  if ((tid & 31) == 0 && local_valid) atomicAdd(&s_valid, local_valid);
  __syncthreads();
  const int n_valid = s_valid;
  const int m = min(k, n_valid);

  unsigned long long thresh = 1ull;   // keep every non-zero key
  if (n_valid > k) {
    // MSB-first radix select of the k-th largest key.
    if (tid == 0) { s_prefix = 0ull; s_need = k; }
    for (int shift = 56; shift >= 0; shift -= 8) {
      for (int i = tid; i < TOPK_WARPS * 256; i += TOPK_THREADS) (&hist[0][0])[i] = 0u;
      __syncthreads();
      if (s_done) break;
      const unsigned long long prefix = s_prefix;
      const unsigned long long himask = (shift == 56) ? 0ull : (~0ull << (shift + 8));
      for (int i = tid; i < n; i += TOPK_THREADS) {
        const unsigned long long key = keys[i];
        if (key != 0ull && (key & himask) == prefix) atomicAdd(&hist[warp][(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 256) {
        unsigned s = 0;
        for (int wv = 0; wv < TOPK_WARPS; ++wv) s += hist[wv][tid];
        total[tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        int need = s_need;
        int d = 255;
        for (; d > 0; --d) {
          if ((int)total[d] >= need) break;
          need -= total[d];
        }
        s_prefix = prefix | ((unsigned long long)d << shift);
        s_need = need;
        if (shift == 0 || (int)total[d] == need) s_done = 1;   // bucket taken whole: low bits are free
      }
      __syncthreads();
    }
    __syncthreads();
    thresh = s_prefix;   // every key >= thresh (in the selected high bits) survives; exactly k of them
  }

  // compaction of survivors into shared memory (order irrelevant, sorted next)
  for (int base = 0; base < n; base += TOPK_THREADS) {
    const int i = base + tid;
    const unsigned long long key = (i < n) ? keys[i] : 0ull;
    const bool take = key != 0ull && key >= thresh;
    const unsigned ballot = __ballot_sync(0xffffffffu, take);
    int wbase = 0;
    if ((tid & 31) == 0 && ballot) wbase = atomicAdd(&s_count, __popc(ballot));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (take) sbuf[wbase + __popc(ballot & ((1u << (tid & 31)) - 1u))] = key;
  }
  __syncthreads();
  for (int i = s_count + tid; i < m_pow2; i += TOPK_THREADS) sbuf[i] = 0ull;   // pad with the smallest key
  bitonic_sort_desc(sbuf, m_pow2);

  const BoxI16* boxes = boxes_all + (size_t)img * n;
  const float* cls = cls_all + (size_t)img * n;
  for (int r = tid; r < k; r += TOPK_THREADS) {
    const size_t o = (size_t)img * k + r;
    if (r < m) {
      const int idx = (int)(unsigned)(sbuf[r] & 0xffffffffull);
      out_boxes[o] = boxes[idx];
      out_scores[o] = __ldg(cls + idx);
      out_index[o] = idx;
    } else {
      out_boxes[o] = BoxI16{0, 0, 0, 0};
      out_scores[o] = 0.0f;
      out_index[o] = -1;
    }
  }
  if (tid == 0) out_count[img] = m;
}

int launch_decode_topk(frcnn_handle* h, cudaStream_t stream, const float* regr, const float* cls,
                       const AnchorTable& tab, int rows, int cols, int k, int batch,
                       int16_t* out_boxes, float* out_scores, int32_t* out_index,
                       int32_t* out_count, float* dense_boxes) {
  const int n = rows * cols * tab.n;
  int m_pow2 = 1;
  while (m_pow2 < k) m_pow2 <<= 1;
  if (m_pow2 > n) { int p = 1; while (p < n) p <<= 1; m_pow2 = p < m_pow2 ? p : m_pow2; }
  const size_t smem = (size_t)m_pow2 * sizeof(unsigned long long);
  if (smem + 40 * 1024 > (size_t)h->max_smem_optin)
    return fail(h, FRCNN_ERR_UNSUPPORTED, "decode_topk: k too large for the shared-memory sort%s%s");
  const size_t key_bytes = align_up((size_t)batch * n * sizeof(unsigned long long), 256);
  const size_t box_bytes = align_up((size_t)batch * n * sizeof(BoxI16), 256);
  void* ws = nullptr;
  int rc = arena_get(h, stream, key_bytes + box_bytes, &ws);
  if (rc) return rc;
  auto* keys = reinterpret_cast<unsigned long long*>(ws);
  auto* boxes = reinterpret_cast<BoxI16*>(reinterpret_cast<char*>(ws) + key_bytes);

  dim3 grid((n + 255) / 256, batch);
  decode_kernel<<<grid, 256, 0, stream>>>(regr, cls, tab, rows, cols, n, keys, boxes,
                                         reinterpret_cast<float4*>(dense_boxes));
  FRCNN_LAUNCH_CHECK(h, "decode_kernel");
  FRCNN_CUDA(h, cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_kernel<<<batch, TOPK_THREADS, smem, stream>>>(keys, boxes, cls, n, k, m_pow2,
                                                     reinterpret_cast<BoxI16*>(out_boxes), out_scores,
                                                     out_index, out_count);
  FRCNN_LAUNCH_CHECK(h, "topk_kernel");
  return FRCNN_OK;
}

}  // namespace frcnn
