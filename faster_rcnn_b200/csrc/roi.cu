// K-d: RoI layer (custom_layers.py:35-56) forward and backward, channels-last.
//
// A tiny pre-kernel turns the RoIs of a launch into tap tables (all int<->float conversions and
// divisions of the layer, once per (RoI, output index)); the streaming kernels only index them.
//
// Forward (both modes): one CTA per (RoI, 1024-channel block, image); a thread owns four
// consecutive channels (128-bit loads/stores), walks the PxP outputs and streams them out with
// evict-first stores.  The feature map (9.8 MB at 38x63x1024) stays L2-resident, the P*P*C
// outputs (401 MB at N=2000) are the HBM stream.
//
// Backward, resize mode: cell-stationary gather, one warp per dX cell accumulating in registers.
// Backward, max mode: spatial-tile ownership in shared memory with per-tile work lists.
// Both walk RoIs in index order and, inside a RoI, bins in (ph, pw, tap) order, so every addition
// into a given dX element happens in one fixed order: no atomics, bit-reproducible run to run.
// (oracle/roi_oracle.py sums each RoI into a private crop first, like TF's slice-gradient +
// AddN, so resize-mode gradients agree to float32 round-off, not bit for bit; max mode is exact.)
//
// RESIZE mode = TF-1.3 legacy bilinear (align_corners=False, no half-pixel offset):
//   scale = in/float(out); src = i*scale; lo = (int)src; hi = min(lo+1, in-1); lerp = src-lo
//   top = tl + (tr-tl)*lx; bottom = bl + (br-bl)*lx; out = top + (bottom-top)*ly
// MAX mode: bin rows y1+floor(ph*h/P) .. y1+ceil((ph+1)*h/P)-1, first maximum in row-major scan.
// This TU is compiled with -fmad=false so that a*b+c keeps two roundings like the CPU oracle.
#include "common.cuh"

namespace frcnn {

struct Crop { int x1, y1, w, h; };   // clipped to the map; w,h <= 0 means empty

__device__ __forceinline__ Crop load_crop(const void* rois, int dtype, size_t idx, int W, int H) {
  int x1, y1, x2, y2;
  if (dtype == FRCNN_ROI_I16) {
    const short* p = reinterpret_cast<const short*>(rois) + idx * 4;
    x1 = p[0]; y1 = p[1]; x2 = p[2]; y2 = p[3];
  } else if (dtype == FRCNN_ROI_I32) {
    const int* p = reinterpret_cast<const int*>(rois) + idx * 4;
    x1 = p[0]; y1 = p[1]; x2 = p[2]; y2 = p[3];
  } else {
    const float* p = reinterpret_cast<const float*>(rois) + idx * 4;
    x1 = (int)p[0]; y1 = (int)p[1]; x2 = (int)p[2]; y2 = (int)p[3];   // K.cast(.., 'int32') truncates
  }
  x1 = max(x1, 0); y1 = max(y1, 0); x2 = min(x2, W); y2 = min(y2, H);
  return Crop{x1, y1, x2 - x1, y2 - y1};
}

struct Tap { int lo, hi; float lerp; };
__device__ __forceinline__ Tap axis_tap(int i, float scale, int in_size) {
  const float src = (float)i * scale;
  Tap t;
  t.lo = (int)src;
  t.hi = min(t.lo + 1, in_size - 1);
  t.lerp = src - (float)t.lo;
  return t;
}

__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
  return make_float4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t);
}

__device__ __forceinline__ float4 scale4(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// ---------------------------------------------------------------------------------------
// per-launch RoI tables (tiny pre-kernel): the crop of every RoI and, per output index p, the
// source taps (resize) or bin bounds (max) of both axes in ABSOLUTE map coordinates.  All integer
// <-> float conversions and divisions of the layer happen here, once per (RoI, p), instead of
// once per (RoI, p, channel block) in the streaming kernels (they run on the 16-lane XU pipe).
//   crops[roi]      = (x1, y1, w, h)
//   taps[roi*P + p] = resize: (ylo | yhi << 16, bits(ylerp), xlo | xhi << 16, bits(xlerp))
//                     max:    (ya  | yb  << 16, 0,           xa  | xb  << 16, 0)   [a, b) bounds
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
roi_table_kernel(const void* __restrict__ rois, int dtype, int n_total, int W, int H, int P,
                 int4* __restrict__ crops, int4* __restrict__ taps) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= n_total * (P + 1)) return;
  const int roi = idx / (P + 1), p = idx - roi * (P + 1);
  const Crop k = load_crop(rois, dtype, (size_t)roi, W, H);
  if (p == P) {
    crops[roi] = make_int4(k.x1, k.y1, k.w, k.h);
    return;
  }
  int4 rec = make_int4(0, 0, 0, 0);
  if (k.w > 0 && k.h > 0) {
    if (MODE == FRCNN_ROI_RESIZE) {
      const Tap ty = axis_tap(p, (float)k.h / (float)P, k.h), tx = axis_tap(p, (float)k.w / (float)P, k.w);
      rec = make_int4((k.y1 + ty.lo) | ((k.y1 + ty.hi) << 16), __float_as_int(ty.lerp),
                      (k.x1 + tx.lo) | ((k.x1 + tx.hi) << 16), __float_as_int(tx.lerp));
    } else {
      const int ya = k.y1 + (p * k.h) / P, yb = k.y1 + ((p + 1) * k.h + P - 1) / P;
      const int xa = k.x1 + (p * k.w) / P, xb = k.x1 + ((p + 1) * k.w + P - 1) / P;
      rec = make_int4(ya | (yb << 16), 0, xa | (xb << 16), 0);
    }
  }
  taps[(size_t)roi * P + p] = rec;
}

constexpr int ROI_FWD_THREADS = 256;
constexpr int ROI_MAX_TABLE_P = 32;     // table-driven kernels support pool sizes up to 32

// Forward: one CTA per (RoI, 1024-channel block, image); a thread owns four consecutive channels.
template <int MODE>
__global__ void __launch_bounds__(ROI_FWD_THREADS)
roi_fwd_kernel(const float* __restrict__ feat, int H, int W, int C, const int4* __restrict__ crops,
               const int4* __restrict__ taps, int N, int P, float* __restrict__ out, int* __restrict__ argmax) {
  __shared__ int4 s_tap[ROI_MAX_TABLE_P];
  __shared__ int4 s_crop;
  const int r = blockIdx.x, img = blockIdx.z;
  const size_t roi = (size_t)img * N + r;
  if (threadIdx.x < P) s_tap[threadIdx.x] = taps[roi * P + threadIdx.x];
  if (threadIdx.x == 32) s_crop = crops[roi];
  __syncthreads();
  const int c = (blockIdx.y * ROI_FWD_THREADS + threadIdx.x) * 4;
  if (c >= C) return;
  const float* f = feat + (size_t)img * H * W * C + c;
  const size_t obase = (roi * P * P) * C + c;
  if (s_crop.z <= 0 || s_crop.w <= 0) {   // TF would raise on an empty crop; we emit zeros
    for (int b = 0; b < P * P; ++b) {
      st_cs_f4(out + obase + (size_t)b * C, make_float4(0.f, 0.f, 0.f, 0.f));
      if (MODE == FRCNN_ROI_MAX) st_cs_i4(argmax + obase + (size_t)b * C, make_int4(0, 0, 0, 0));
    }
    return;
  }
  const size_t row_stride = (size_t)W * C;
  if (MODE == FRCNN_ROI_RESIZE) {
    for (int ph = 0; ph < P; ++ph) {
      const int4 ty = s_tap[ph];
      const float* row_lo = f + (size_t)(ty.x & 0xffff) * row_stride;
      const float* row_hi = f + (size_t)(ty.x >> 16) * row_stride;
      const float ly = __int_as_float(ty.y);
      float* o = out + obase + (size_t)ph * P * C;
      // Measured and rejected (profiles/roi_fwd_r01.md): issuing the tap loads of 2/4/7 outputs ahead
      // (no gain: the kernel is bound by L1/TEX + DRAM-write throughput, not load latency) and keeping
      // the last two source columns in registers (fewer loads, but the extra registers cost more
      // occupancy than the loads saved).
      for (int pw = 0; pw < P; ++pw) {
        const int4 tx = s_tap[pw];
        const size_t xl = (size_t)(tx.z & 0xffff) * C, xh = (size_t)(tx.z >> 16) * C;
        const float lx = __int_as_float(tx.w);
        const float4 tl = ldg_f4(row_lo + xl), tr = ldg_f4(row_lo + xh);
        const float4 bl = ldg_f4(row_hi + xl), br = ldg_f4(row_hi + xh);
        const float4 top = lerp4(tl, tr, lx), bot = lerp4(bl, br, lx);
        st_cs_f4(o + (size_t)pw * C, lerp4(top, bot, ly));
      }
    }
  } else {
    for (int ph = 0; ph < P; ++ph) {
      const int ya = s_tap[ph].x & 0xffff, yb = s_tap[ph].x >> 16;
      for (int pw = 0; pw < P; ++pw) {
        const int xa = s_tap[pw].z & 0xffff, xb = s_tap[pw].z >> 16;
        float4 best = ldg_f4(f + ((size_t)ya * W + xa) * C);
        int4 arg = make_int4(ya * W + xa, ya * W + xa, ya * W + xa, ya * W + xa);
        for (int y = ya; y < yb; ++y) {
          for (int x = xa; x < xb; ++x) {
            const float4 v = ldg_f4(f + ((size_t)y * W + x) * C);
            const int cell = y * W + x;
            if (v.x > best.x) { best.x = v.x; arg.x = cell; }
            if (v.y > best.y) { best.y = v.y; arg.y = cell; }
            if (v.z > best.z) { best.z = v.z; arg.z = cell; }
            if (v.w > best.w) { best.w = v.w; arg.w = cell; }
          }
        }
        st_cs_f4(out + obase + (size_t)(ph * P + pw) * C, best);
        st_cs_i4(argmax + obase + (size_t)(ph * P + pw) * C, arg);
      }
    }
  }
}

// scalar-channel fallback for C % 4 != 0 (not a performance path)
template <int MODE>
__global__ void roi_fwd_scalar_kernel(const float* __restrict__ feat, int H, int W, int C,
                                      const void* __restrict__ rois, int dtype, int N, int P,
                                      float* __restrict__ out, int* __restrict__ argmax) {
  const int r = blockIdx.x, img = blockIdx.z;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const Crop k = load_crop(rois, dtype, (size_t)img * N + r, W, H);
  const float* f = feat + (size_t)img * H * W * C + c;
  const size_t obase = (((size_t)img * N + r) * P * P) * C + c;
  if (k.w <= 0 || k.h <= 0) {
    for (int b = 0; b < P * P; ++b) { out[obase + (size_t)b * C] = 0.f; if (MODE == FRCNN_ROI_MAX) argmax[obase + (size_t)b * C] = 0; }
    return;
  }
  const float ys = (float)k.h / (float)P, xs = (float)k.w / (float)P;
  for (int ph = 0; ph < P; ++ph) {
    for (int pw = 0; pw < P; ++pw) {
      const size_t o = obase + (size_t)(ph * P + pw) * C;
      if (MODE == FRCNN_ROI_RESIZE) {
        const Tap ty = axis_tap(ph, ys, k.h), tx = axis_tap(pw, xs, k.w);
        const float tl = f[((size_t)(k.y1 + ty.lo) * W + k.x1 + tx.lo) * C], tr = f[((size_t)(k.y1 + ty.lo) * W + k.x1 + tx.hi) * C];
        const float bl = f[((size_t)(k.y1 + ty.hi) * W + k.x1 + tx.lo) * C], br = f[((size_t)(k.y1 + ty.hi) * W + k.x1 + tx.hi) * C];
        const float top = tl + (tr - tl) * tx.lerp, bot = bl + (br - bl) * tx.lerp;
        out[o] = top + (bot - top) * ty.lerp;
      } else {
        const int ya = k.y1 + (ph * k.h) / P, yb = k.y1 + ((ph + 1) * k.h + P - 1) / P;
        const int xa = k.x1 + (pw * k.w) / P, xb = k.x1 + ((pw + 1) * k.w + P - 1) / P;
        float best = f[((size_t)ya * W + xa) * C];
        int arg = ya * W + xa;
        for (int y = ya; y < yb; ++y)
          for (int x = xa; x < xb; ++x) {
            const float v = f[((size_t)y * W + x) * C];
            if (v > best) { best = v; arg = y * W + x; }
          }
        out[o] = best;
        argmax[o] = arg;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward: spatial-tile ownership, work lists, float4 channel columns
//
// A CTA owns a 4x8-cell tile of dX for 256 channels in shared memory; a thread owns one float4
// (resize) / four strided (max) channel columns of it, so no two threads ever touch the same
// accumulator: no atomics, no barriers in the accumulation loop, one fixed summation order.
// RoIs are taken in chunks: one thread per RoI finds the contiguous (ph, pw) ranges whose taps /
// bins can touch the tile, an ordered block scan turns the counts into offsets, and the
// (roi, ph, pw) work items land in a shared list in ascending (roi, ph, pw) order.  Then every
// thread streams through the list, issuing the dY loads of several items before consuming them.
// ---------------------------------------------------------------------------------------
constexpr int BT_H = 4, BT_W = 8, BT_CELLS = BT_H * BT_W;
constexpr int BT_THREADS = 64;
constexpr int BT_CH = BT_THREADS * 4;          // channels per CTA
constexpr int BT_LIST = 1568;                  // work items per RoI chunk (32 RoIs x 49 bins at P = 7)
constexpr int BT_UNROLL = 4;                   // dY loads in flight per thread

__device__ __forceinline__ void add4(float4* p, float4 v) {
  float4 a = *p;
  a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  *p = a;
}

template <int MODE>
__global__ void __launch_bounds__(BT_THREADS)
roi_bwd_tile_kernel(const float* __restrict__ gout, const void* __restrict__ rois, int dtype,
                    const int* __restrict__ argmax, int H, int W, int C, int N, int P, int tiles_x, int chunk,
                    float* __restrict__ gfeat) {
  __shared__ __align__(16) float s_acc[BT_CELLS * BT_CH];
  __shared__ unsigned s_list[BT_LIST];
  __shared__ int4 s_crop[BT_THREADS];
  __shared__ float2 s_scale[BT_THREADS];
  __shared__ int s_wtot[BT_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, img = blockIdx.z;
  const int ty0 = (tile / tiles_x) * BT_H, tx0 = (tile % tiles_x) * BT_W;
  const int th = min(BT_H, H - ty0), tw = min(BT_W, W - tx0);
  const int cbase = blockIdx.y * BT_CH;
  for (int i = tid; i < BT_CELLS * BT_CH; i += BT_THREADS) s_acc[i] = 0.f;

  const size_t img_off = (size_t)img * N * P * P * C;
  const float* g_img = gout + img_off;
  const int* a_img = (MODE == FRCNN_ROI_MAX) ? argmax + img_off : nullptr;
  const int c4 = cbase + 4 * tid;                 // resize: four consecutive channels
  const bool live4 = c4 < C;
  float4* acc4 = reinterpret_cast<float4*>(s_acc) + tid;      // cell stride = BT_THREADS float4

  for (int base = 0; base < N; base += chunk) {
    // ---- build the work list of this RoI chunk (thread t <-> RoI base + t) ----
    int pa = 0, pb = 0, qa = 0, qb = 0;
    if (tid < chunk && base + tid < N) {
      const Crop k = load_crop(rois, dtype, (size_t)img * N + base + tid, W, H);
      s_crop[tid] = make_int4(k.x1, k.y1, k.w, k.h);
      if (k.w > 0 && k.h > 0 && k.x1 < tx0 + tw && k.x1 + k.w > tx0 && k.y1 < ty0 + th && k.y1 + k.h > ty0) {
        const float ys = (float)k.h / (float)P, xs = (float)k.w / (float)P;
        s_scale[tid] = make_float2(ys, xs);
        pa = P; qa = P;
        for (int p = 0; p < P; ++p) {
          bool hy, hx;
          if (MODE == FRCNN_ROI_RESIZE) {
            const Tap t = axis_tap(p, ys, k.h), u = axis_tap(p, xs, k.w);
            hy = (unsigned)(k.y1 + t.lo - ty0) < (unsigned)th || (unsigned)(k.y1 + t.hi - ty0) < (unsigned)th;
            hx = (unsigned)(k.x1 + u.lo - tx0) < (unsigned)tw || (unsigned)(k.x1 + u.hi - tx0) < (unsigned)tw;
          } else {
            hy = k.y1 + (p * k.h) / P < ty0 + th && k.y1 + ((p + 1) * k.h + P - 1) / P > ty0;
            hx = k.x1 + (p * k.w) / P < tx0 + tw && k.x1 + ((p + 1) * k.w + P - 1) / P > tx0;
          }
          if (hy) { pa = min(pa, p); pb = p + 1; }
          if (hx) { qa = min(qa, p); qb = p + 1; }
        }
        if (pb == 0 || qb == 0) { pa = pb = qa = qb = 0; }
      }
    }
    const int cnt = (pb - pa) * (qb - qa);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    int off = incl - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < BT_THREADS / 32; ++w) {
      if (w < warp) off += s_wtot[w];
      total += s_wtot[w];
    }
    for (int ph = pa; ph < pb; ++ph)
      for (int pw = qa; pw < qb; ++pw) s_list[off++] = ((unsigned)tid << 16) | ((unsigned)ph << 8) | (unsigned)pw;
    __syncthreads();

    // ---- consume: every thread walks the whole list for its own channel columns ----
    if (MODE == FRCNN_ROI_RESIZE) {
      for (int i0 = 0; i0 < total; i0 += BT_UNROLL) {
        float4 g[BT_UNROLL];
        unsigned e[BT_UNROLL];
#pragma unroll
        for (int u = 0; u < BT_UNROLL; ++u) {
          g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          e[u] = 0u;
          if (i0 + u < total) {
            e[u] = s_list[i0 + u];
            const int r = base + (int)(e[u] >> 16), ph = (e[u] >> 8) & 255, pw = e[u] & 255;
            if (live4) g[u] = ldg_f4(g_img + ((size_t)(r * P + ph) * P + pw) * C + c4);
          }
        }
#pragma unroll
        for (int u = 0; u < BT_UNROLL; ++u) {
          if (i0 + u < total) {
            const int t = (int)(e[u] >> 16), ph = (e[u] >> 8) & 255, pw = e[u] & 255;
            const int4 k = s_crop[t];
            const float2 sc = s_scale[t];
            const Tap ty = axis_tap(ph, sc.x, k.w), tx = axis_tap(pw, sc.y, k.z);
            const int ylo = k.y + ty.lo - ty0, yhi = k.y + ty.hi - ty0;
            const int xlo = k.x + tx.lo - tx0, xhi = k.x + tx.hi - tx0;
            const bool rlo = (unsigned)ylo < (unsigned)th, rhi = (unsigned)yhi < (unsigned)th;
            const bool clo = (unsigned)xlo < (unsigned)tw, chi = (unsigned)xhi < (unsigned)tw;
            // order TL, TR, BL, BR; weight product (g*wy)*wx as in ResizeBilinearGrad
            const float4 gy0 = scale4(g[u], 1.0f - ty.lerp), gy1 = scale4(g[u], ty.lerp);
            const float wx0 = 1.0f - tx.lerp, wx1 = tx.lerp;
            if (rlo && clo) add4(acc4 + (ylo * BT_W + xlo) * BT_THREADS, scale4(gy0, wx0));
            if (rlo && chi) add4(acc4 + (ylo * BT_W + xhi) * BT_THREADS, scale4(gy0, wx1));
            if (rhi && clo) add4(acc4 + (yhi * BT_W + xlo) * BT_THREADS, scale4(gy1, wx0));
            if (rhi && chi) add4(acc4 + (yhi * BT_W + xhi) * BT_THREADS, scale4(gy1, wx1));
          }
        }
      }
    } else {
      // max mode: the arg-max cell differs per channel; channels are strided (c = cbase + j*64 + tid)
      // so that both the global loads and the shared-memory columns are conflict-free.
      const int first = ty0 * W + tx0;
      for (int i = 0; i < total; ++i) {
        const unsigned e = s_list[i];
        const int r = base + (int)(e >> 16), ph = (e >> 8) & 255, pw = e & 255;
        const size_t o = ((size_t)(r * P + ph) * P + pw) * C;
        int a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cbase + j * BT_THREADS + tid;
          a[j] = (c < C) ? __ldg(a_img + o + c) : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cbase + j * BT_THREADS + tid;
          int cell = -1;
#pragma unroll
          for (int y = 0; y < BT_H; ++y) {
            const int d = a[j] - (first + y * W);
            if (y < th && (unsigned)d < (unsigned)tw) cell = y * BT_W + d;
          }
          if (cell >= 0 && a[j] >= 0) s_acc[cell * BT_CH + j * BT_THREADS + tid] += __ldg(g_img + o + c);
        }
      }
    }
    __syncthreads();      // the list and crop tables are rewritten by the next chunk
  }

  float* dst = gfeat + (size_t)img * H * W * C;
  if (MODE == FRCNN_ROI_RESIZE) {
    if (!live4) return;
    for (int y = 0; y < th; ++y)
      for (int x = 0; x < tw; ++x)
        *reinterpret_cast<float4*>(dst + ((size_t)(ty0 + y) * W + tx0 + x) * C + c4) = acc4[(y * BT_W + x) * BT_THREADS];
  } else {
    for (int y = 0; y < th; ++y)
      for (int x = 0; x < tw; ++x)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cbase + j * BT_THREADS + tid;
          if (c < C) dst[((size_t)(ty0 + y) * W + tx0 + x) * C + c] = s_acc[(y * BT_W + x) * BT_CH + j * BT_THREADS + tid];
        }
  }
}

// ---------------------------------------------------------------------------------------
// backward, resize mode: cell-stationary gather, one warp per dX cell
//
// A warp owns one feature-map cell (y, x) for up to 1024 channels (lane l holds float4 channel
// blocks l, l+32, ...) and accumulates in registers -- no shared memory, no atomics, one store.
// It walks the RoIs 32 at a time (lane <-> RoI containment test, ballot), and for every RoI whose
// crop contains the cell the lanes compute the P y-taps and P x-taps in parallel (lane <-> tap
// index); two ballots give the bins whose taps land on this cell.  Each matching (ph, pw) bin is
// one coalesced dY row read (512 B per channel block) scaled by (g*wy)*wx.  Contributions are
// added in ascending (roi, ph, pw, TL/TR/BL/BR) order, so the result is bit-reproducible.
// dY rows are shared by the <= 4 neighbouring cells they touch; neighbouring cells sit in the
// same CTA (x direction) or the same wave (y direction), so the re-reads hit L1 / L2, not DRAM.
// ---------------------------------------------------------------------------------------
constexpr int CW_WARPS = 8;

template <int CPB>
__global__ void __launch_bounds__(CW_WARPS * 32)
roi_bwd_resize_cell_kernel(const float* __restrict__ gout, const int4* __restrict__ crops,
                           const int4* __restrict__ taps, int H, int W, int C, int N, int P,
                           float* __restrict__ gfeat) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cell = blockIdx.x * CW_WARPS + warp;
  if (cell >= H * W) return;                        // warp-uniform; the kernel has no block-level barrier
  const int img = blockIdx.z;
  const int y = cell / W, x = cell - y * W;
  const int cbase = blockIdx.y * (CPB * 128) + 4 * lane;
  const float* g_img = gout + (size_t)img * N * P * P * C;
  const int4* crop_img = crops + (size_t)img * N;
  const int4* tap_img = taps + (size_t)img * N * P;

  float4 acc[CPB];
#pragma unroll
  for (int j = 0; j < CPB; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int w0 = 0; w0 < N; w0 += 32) {
    const int r = w0 + lane;
    bool inside = false;
    if (r < N) {
      const int4 k = __ldg(crop_img + r);
      inside = k.z > 0 && k.w > 0 && x >= k.x && x < k.x + k.z && y >= k.y && y < k.y + k.w;
    }
    unsigned m = __ballot_sync(0xffffffffu, inside);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      // lane p looks at tap p of both axes; code bit0 = "lo tap is this cell", bit1 = "hi tap is this cell"
      int ycode = 0, xcode = 0;
      float ly = 0.f, lx = 0.f;
      if (lane < P) {
        const int4 t = __ldg(tap_img + (size_t)(w0 + b) * P + lane);
        ycode = ((t.x & 0xffff) == y ? 1 : 0) | ((t.x >> 16) == y ? 2 : 0);
        xcode = ((t.z & 0xffff) == x ? 1 : 0) | ((t.z >> 16) == x ? 2 : 0);
        ly = __int_as_float(t.y);
        lx = __int_as_float(t.w);
      }
      unsigned my = __ballot_sync(0xffffffffu, ycode != 0);
      const unsigned mx = __ballot_sync(0xffffffffu, xcode != 0);
      if (mx == 0u) continue;
      const size_t roi_row = (size_t)(w0 + b) * P;
      while (my) {
        const int ph = __ffs(my) - 1;
        my &= my - 1;
        const int yc = __shfl_sync(0xffffffffu, ycode, ph);
        const float wy1 = __shfl_sync(0xffffffffu, ly, ph), wy0 = 1.0f - wy1;
        unsigned mxx = mx;
        while (mxx) {
          const int pw = __ffs(mxx) - 1;
          mxx &= mxx - 1;
          const int xc = __shfl_sync(0xffffffffu, xcode, pw);
          const float wx1 = __shfl_sync(0xffffffffu, lx, pw), wx0 = 1.0f - wx1;
          const float* row = g_img + ((roi_row + ph) * P + pw) * C + cbase;
          float4 g[CPB];
#pragma unroll
          for (int j = 0; j < CPB; ++j)
            g[j] = (cbase + j * 128 < C) ? ldg_f4(row + j * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < CPB; ++j) {
            // order TL, TR, BL, BR; weight product (g*wy)*wx as in ResizeBilinearGrad
            if (yc & 1) {
              const float4 t = scale4(g[j], wy0);
              if (xc & 1) { const float4 v = scale4(t, wx0); acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w; }
              if (xc & 2) { const float4 v = scale4(t, wx1); acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w; }
            }
            if (yc & 2) {
              const float4 t = scale4(g[j], wy1);
              if (xc & 1) { const float4 v = scale4(t, wx0); acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w; }
              if (xc & 2) { const float4 v = scale4(t, wx1); acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w; }
            }
          }
        }
      }
    }
  }
  float* dst = gfeat + ((size_t)img * H * W + cell) * C + cbase;
#pragma unroll
  for (int j = 0; j < CPB; ++j)
    if (cbase + j * 128 < C) *reinterpret_cast<float4*>(dst + j * 128) = acc[j];
}

// ---------------------------------------------------------------------------------------
// backward, scalar-channel fallback (C % 4 != 0 or pool sizes beyond the work-list capacity): 8x8 tile,
// one channel per thread, every CTA scans all RoIs.  Not a performance path.
// ---------------------------------------------------------------------------------------
constexpr int BWD_TILE = 8;          // 8x8 cells
constexpr int BWD_CH = 512;          // channels per CTA = threads per CTA

template <int MODE>
__global__ void __launch_bounds__(BWD_CH, 1)
roi_bwd_scalar_kernel(const float* __restrict__ gout, const void* __restrict__ rois, int dtype,
               const int* __restrict__ argmax, int H, int W, int C, int N, int P, int tiles_x,
               float* __restrict__ gfeat) {
  extern __shared__ float acc[];     // [BWD_TILE*BWD_TILE][BWD_CH]
  const int tile = blockIdx.x, img = blockIdx.z;
  const int ty0 = (tile / tiles_x) * BWD_TILE, tx0 = (tile % tiles_x) * BWD_TILE;
  const int ty1 = min(ty0 + BWD_TILE, H), tx1 = min(tx0 + BWD_TILE, W);
  const int c = blockIdx.y * BWD_CH + threadIdx.x;
  const bool live = c < C;
  for (int i = 0; i < BWD_TILE * BWD_TILE; ++i) acc[i * BWD_CH + threadIdx.x] = 0.f;

  const float* g_img = gout + (size_t)img * N * P * P * C + c;
  const int* a_img = (MODE == FRCNN_ROI_MAX) ? argmax + (size_t)img * N * P * P * C + c : nullptr;

  for (int r = 0; r < N; ++r) {
    const Crop k = load_crop(rois, dtype, (size_t)img * N + r, W, H);
    if (k.w <= 0 || k.h <= 0) continue;
    if (k.x1 >= tx1 || k.x1 + k.w <= tx0 || k.y1 >= ty1 || k.y1 + k.h <= ty0) continue;   // CTA-uniform
    const float* g_roi = g_img + (size_t)r * P * P * C;
    if (MODE == FRCNN_ROI_RESIZE) {
      const float ys = (float)k.h / (float)P, xs = (float)k.w / (float)P;
      for (int ph = 0; ph < P; ++ph) {
        const Tap ty = axis_tap(ph, ys, k.h);
        const int ylo = k.y1 + ty.lo, yhi = k.y1 + ty.hi;
        const bool rlo = ylo >= ty0 && ylo < ty1, rhi = yhi >= ty0 && yhi < ty1;
        if (!rlo && !rhi) continue;
        const float wy1 = ty.lerp, wy0 = 1.0f - ty.lerp;
        for (int pw = 0; pw < P; ++pw) {
          const Tap tx = axis_tap(pw, xs, k.w);
          const int xlo = k.x1 + tx.lo, xhi = k.x1 + tx.hi;
          const bool clo = xlo >= tx0 && xlo < tx1, chi = xhi >= tx0 && xhi < tx1;
          if (!clo && !chi) continue;
          const float g = live ? __ldg(g_roi + (size_t)(ph * P + pw) * C) : 0.f;
          const float wx1 = tx.lerp, wx0 = 1.0f - tx.lerp;
          // order TL, TR, BL, BR; weight product (g*wy)*wx as in ResizeBilinearGrad
          if (rlo && clo) acc[((ylo - ty0) * BWD_TILE + (xlo - tx0)) * BWD_CH + threadIdx.x] += g * wy0 * wx0;
          if (rlo && chi) acc[((ylo - ty0) * BWD_TILE + (xhi - tx0)) * BWD_CH + threadIdx.x] += g * wy0 * wx1;
          if (rhi && clo) acc[((yhi - ty0) * BWD_TILE + (xlo - tx0)) * BWD_CH + threadIdx.x] += g * wy1 * wx0;
          if (rhi && chi) acc[((yhi - ty0) * BWD_TILE + (xhi - tx0)) * BWD_CH + threadIdx.x] += g * wy1 * wx1;
        }
      }
    } else {
      for (int ph = 0; ph < P; ++ph) {
        const int ya = k.y1 + (ph * k.h) / P, yb = k.y1 + ((ph + 1) * k.h + P - 1) / P;
        if (ya >= ty1 || yb <= ty0) continue;
        for (int pw = 0; pw < P; ++pw) {
          const int xa = k.x1 + (pw * k.w) / P, xb = k.x1 + ((pw + 1) * k.w + P - 1) / P;
          if (xa >= tx1 || xb <= tx0) continue;
          if (!live) continue;
          const size_t o = (size_t)r * P * P * C + (size_t)(ph * P + pw) * C;
          const int cell = __ldg(a_img + o);
          const int ay = cell / W, ax = cell - ay * W;
          if (ay >= ty0 && ay < ty1 && ax >= tx0 && ax < tx1)
            acc[((ay - ty0) * BWD_TILE + (ax - tx0)) * BWD_CH + threadIdx.x] += __ldg(g_img + o);
        }
      }
    }
  }
  if (!live) return;
  float* dst = gfeat + (size_t)img * H * W * C + c;
  for (int y = ty0; y < ty1; ++y)
    for (int x = tx0; x < tx1; ++x)
      dst[((size_t)y * W + x) * C] = acc[((y - ty0) * BWD_TILE + (x - tx0)) * BWD_CH + threadIdx.x];
}

// builds the per-launch RoI tables in the handle's scratch arena
static int build_tables(frcnn_handle* h, cudaStream_t stream, int mode, const void* rois, int dtype, int n_total,
                        int W, int H, int P, int4** crops, int4** taps) {
  void *pc = nullptr, *pt = nullptr;
  int rc = arena_get(h, stream, (size_t)n_total * sizeof(int4), &pc);
  if (rc) return rc;
  if ((rc = arena_get(h, stream, (size_t)n_total * P * sizeof(int4), &pt))) return rc;
  *crops = static_cast<int4*>(pc);
  *taps = static_cast<int4*>(pt);
  const int blocks = (int)(((size_t)n_total * (P + 1) + 255) / 256);
  if (mode == FRCNN_ROI_RESIZE)
    roi_table_kernel<FRCNN_ROI_RESIZE><<<blocks, 256, 0, stream>>>(rois, dtype, n_total, W, H, P, *crops, *taps);
  else
    roi_table_kernel<FRCNN_ROI_MAX><<<blocks, 256, 0, stream>>>(rois, dtype, n_total, W, H, P, *crops, *taps);
  FRCNN_LAUNCH_CHECK(h, "roi_table_kernel");
  return FRCNN_OK;
}

int launch_roi_fwd(frcnn_handle* h, cudaStream_t stream, int mode, const float* feat, int H, int W, int C,
                   const void* rois, int dtype, int N, int P, int batch, float* out, int32_t* argmax) {
  if (C % 4 == 0 && P <= ROI_MAX_TABLE_P && H < 32768 && W < 32768 && (reinterpret_cast<uintptr_t>(feat) % 16 == 0) &&
      (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
      (mode != FRCNN_ROI_MAX || reinterpret_cast<uintptr_t>(argmax) % 16 == 0)) {
    int4 *crops = nullptr, *taps = nullptr;
    int rc = build_tables(h, stream, mode, rois, dtype, batch * N, W, H, P, &crops, &taps);
    if (rc) return rc;
    dim3 grid(N, (C / 4 + ROI_FWD_THREADS - 1) / ROI_FWD_THREADS, batch);
    if (mode == FRCNN_ROI_RESIZE)
      roi_fwd_kernel<FRCNN_ROI_RESIZE><<<grid, ROI_FWD_THREADS, 0, stream>>>(feat, H, W, C, crops, taps, N, P, out, argmax);
    else
      roi_fwd_kernel<FRCNN_ROI_MAX><<<grid, ROI_FWD_THREADS, 0, stream>>>(feat, H, W, C, crops, taps, N, P, out, argmax);
  } else {
    dim3 grid(N, (C + 127) / 128, batch);
    if (mode == FRCNN_ROI_RESIZE)
      roi_fwd_scalar_kernel<FRCNN_ROI_RESIZE><<<grid, 128, 0, stream>>>(feat, H, W, C, rois, dtype, N, P, out, argmax);
    else
      roi_fwd_scalar_kernel<FRCNN_ROI_MAX><<<grid, 128, 0, stream>>>(feat, H, W, C, rois, dtype, N, P, out, argmax);
  }
  FRCNN_LAUNCH_CHECK(h, "roi_fwd_kernel");
  return FRCNN_OK;
}

int launch_roi_bwd(frcnn_handle* h, cudaStream_t stream, int mode, const float* gout, const void* rois, int dtype,
                   const int32_t* argmax, int H, int W, int C, int N, int P, int batch, float* gfeat) {
  const int chunk = (P * P <= BT_LIST) ? (BT_LIST / (P * P) < BT_THREADS ? BT_LIST / (P * P) : BT_THREADS) : 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(gout) % 16 == 0) && (reinterpret_cast<uintptr_t>(gfeat) % 16 == 0);
  if (mode == FRCNN_ROI_RESIZE && C % 4 == 0 && aligned && P <= ROI_MAX_TABLE_P && H < 32768 && W < 32768) {
    int4 *crops = nullptr, *taps = nullptr;
    int rc = build_tables(h, stream, mode, rois, dtype, batch * N, W, H, P, &crops, &taps);
    if (rc) return rc;
    const int blocks128 = (C + 127) / 128;
    const int cpb = blocks128 >= 8 ? 8 : (blocks128 >= 4 ? 4 : (blocks128 >= 2 ? 2 : 1));
    dim3 grid((H * W + CW_WARPS - 1) / CW_WARPS, (blocks128 + cpb - 1) / cpb, batch);
    if (cpb == 8) roi_bwd_resize_cell_kernel<8><<<grid, CW_WARPS * 32, 0, stream>>>(gout, crops, taps, H, W, C, N, P, gfeat);
    else if (cpb == 4) roi_bwd_resize_cell_kernel<4><<<grid, CW_WARPS * 32, 0, stream>>>(gout, crops, taps, H, W, C, N, P, gfeat);
    else if (cpb == 2) roi_bwd_resize_cell_kernel<2><<<grid, CW_WARPS * 32, 0, stream>>>(gout, crops, taps, H, W, C, N, P, gfeat);
    else roi_bwd_resize_cell_kernel<1><<<grid, CW_WARPS * 32, 0, stream>>>(gout, crops, taps, H, W, C, N, P, gfeat);
    FRCNN_LAUNCH_CHECK(h, "roi_bwd_resize_cell_kernel");
    return FRCNN_OK;
  }
  if (C % 4 == 0 && aligned && chunk > 0 && P <= 255) {
    const int tiles_x = (W + BT_W - 1) / BT_W, tiles_y = (H + BT_H - 1) / BT_H;
    dim3 grid(tiles_x * tiles_y, (C + BT_CH - 1) / BT_CH, batch);
    if (mode == FRCNN_ROI_RESIZE)
      roi_bwd_tile_kernel<FRCNN_ROI_RESIZE><<<grid, BT_THREADS, 0, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, chunk, gfeat);
    else
      roi_bwd_tile_kernel<FRCNN_ROI_MAX><<<grid, BT_THREADS, 0, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, chunk, gfeat);
    FRCNN_LAUNCH_CHECK(h, "roi_bwd_tile_kernel");
    return FRCNN_OK;
  }
  const int tiles_x = (W + BWD_TILE - 1) / BWD_TILE, tiles_y = (H + BWD_TILE - 1) / BWD_TILE;
  dim3 grid(tiles_x * tiles_y, (C + BWD_CH - 1) / BWD_CH, batch);
  const size_t smem = (size_t)BWD_TILE * BWD_TILE * BWD_CH * sizeof(float);
  if (mode == FRCNN_ROI_RESIZE) {
    FRCNN_CUDA(h, cudaFuncSetAttribute(roi_bwd_scalar_kernel<FRCNN_ROI_RESIZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_bwd_scalar_kernel<FRCNN_ROI_RESIZE><<<grid, BWD_CH, smem, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, gfeat);
  } else {
    FRCNN_CUDA(h, cudaFuncSetAttribute(roi_bwd_scalar_kernel<FRCNN_ROI_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_bwd_scalar_kernel<FRCNN_ROI_MAX><<<grid, BWD_CH, smem, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, gfeat);
  }
  FRCNN_LAUNCH_CHECK(h, "roi_bwd_scalar_kernel");
  return FRCNN_OK;
}

}  // namespace frcnn
