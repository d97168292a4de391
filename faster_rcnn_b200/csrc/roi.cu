// K-d: RoI layer (custom_layers.py:35-56) forward and backward, channels-last.
//
// Forward (both modes): one CTA per (RoI, 1024-channel block, image); a thread owns four
// consecutive channels (128-bit loads/stores), walks the PxP outputs and streams them out with
// evict-first stores.  The feature map (9.8 MB at 38x63x1024) stays L2-resident, the P*P*C
// outputs (401 MB at N=2000) are the HBM stream.
//
// Backward (both modes): spatial-tile ownership.  A CTA owns an 8x8-cell tile of dX for a
// 512-channel chunk in shared memory, a thread owns one channel column of it.  RoIs are walked
// in index order, and inside a RoI the taps / bins in (ph, pw, tap) order, so every addition
// into a given dX element happens in one fixed order: no atomics, bit-reproducible, and equal
// to oracle/roi_oracle.py's accumulation order.  dY rows are read coalesced (2 KB per warp-set).
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

constexpr int ROI_FWD_THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(ROI_FWD_THREADS)
roi_fwd_kernel(const float* __restrict__ feat, int H, int W, int C, const void* __restrict__ rois, int dtype,
               int N, int P, float* __restrict__ out, int* __restrict__ argmax) {
  const int r = blockIdx.x, img = blockIdx.z;
  const int c = (blockIdx.y * ROI_FWD_THREADS + threadIdx.x) * 4;
  if (c >= C) return;
  const Crop k = load_crop(rois, dtype, (size_t)img * N + r, W, H);
  const float* f = feat + (size_t)img * H * W * C + c;
  const size_t obase = (((size_t)img * N + r) * P * P) * C + c;
  if (k.w <= 0 || k.h <= 0) {   // TF would raise on an empty crop; we emit zeros
    for (int b = 0; b < P * P; ++b) {
      st_cs_f4(out + obase + (size_t)b * C, make_float4(0.f, 0.f, 0.f, 0.f));
      if (MODE == FRCNN_ROI_MAX) st_cs_i4(argmax + obase + (size_t)b * C, make_int4(0, 0, 0, 0));
    }
    return;
  }
  if (MODE == FRCNN_ROI_RESIZE) {
    const float ys = (float)k.h / (float)P, xs = (float)k.w / (float)P;
    for (int ph = 0; ph < P; ++ph) {
      const Tap ty = axis_tap(ph, ys, k.h);
      const float* row_lo = f + (size_t)(k.y1 + ty.lo) * W * C;
      const float* row_hi = f + (size_t)(k.y1 + ty.hi) * W * C;
      for (int pw = 0; pw < P; ++pw) {
        const Tap tx = axis_tap(pw, xs, k.w);
        const size_t xl = (size_t)(k.x1 + tx.lo) * C, xh = (size_t)(k.x1 + tx.hi) * C;
        const float4 tl = ldg_f4(row_lo + xl), tr = ldg_f4(row_lo + xh);
        const float4 bl = ldg_f4(row_hi + xl), br = ldg_f4(row_hi + xh);
        const float4 top = lerp4(tl, tr, tx.lerp), bot = lerp4(bl, br, tx.lerp);
        st_cs_f4(out + obase + (size_t)(ph * P + pw) * C, lerp4(top, bot, ty.lerp));
      }
    }
  } else {
    for (int ph = 0; ph < P; ++ph) {
      const int ya = k.y1 + (ph * k.h) / P, yb = k.y1 + ((ph + 1) * k.h + P - 1) / P;
      for (int pw = 0; pw < P; ++pw) {
        const int xa = k.x1 + (pw * k.w) / P, xb = k.x1 + ((pw + 1) * k.w + P - 1) / P;
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
// backward: tile ownership
// ---------------------------------------------------------------------------------------
constexpr int BWD_TILE = 8;          // 8x8 cells
constexpr int BWD_CH = 512;          // channels per CTA = threads per CTA

template <int MODE>
__global__ void __launch_bounds__(BWD_CH, 1)
roi_bwd_kernel(const float* __restrict__ gout, const void* __restrict__ rois, int dtype,
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

int launch_roi_fwd(frcnn_handle* h, cudaStream_t stream, int mode, const float* feat, int H, int W, int C,
                   const void* rois, int dtype, int N, int P, int batch, float* out, int32_t* argmax) {
  if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
      (mode != FRCNN_ROI_MAX || reinterpret_cast<uintptr_t>(argmax) % 16 == 0)) {
    dim3 grid(N, (C / 4 + ROI_FWD_THREADS - 1) / ROI_FWD_THREADS, batch);
    if (mode == FRCNN_ROI_RESIZE)
      roi_fwd_kernel<FRCNN_ROI_RESIZE><<<grid, ROI_FWD_THREADS, 0, stream>>>(feat, H, W, C, rois, dtype, N, P, out, argmax);
    else
      roi_fwd_kernel<FRCNN_ROI_MAX><<<grid, ROI_FWD_THREADS, 0, stream>>>(feat, H, W, C, rois, dtype, N, P, out, argmax);
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
  const int tiles_x = (W + BWD_TILE - 1) / BWD_TILE, tiles_y = (H + BWD_TILE - 1) / BWD_TILE;
  dim3 grid(tiles_x * tiles_y, (C + BWD_CH - 1) / BWD_CH, batch);
  const size_t smem = (size_t)BWD_TILE * BWD_TILE * BWD_CH * sizeof(float);
  if (mode == FRCNN_ROI_RESIZE) {
    FRCNN_CUDA(h, cudaFuncSetAttribute(roi_bwd_kernel<FRCNN_ROI_RESIZE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_bwd_kernel<FRCNN_ROI_RESIZE><<<grid, BWD_CH, smem, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, gfeat);
  } else {
    FRCNN_CUDA(h, cudaFuncSetAttribute(roi_bwd_kernel<FRCNN_ROI_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_bwd_kernel<FRCNN_ROI_MAX><<<grid, BWD_CH, smem, stream>>>(gout, rois, dtype, argmax, H, W, C, N, P, tiles_x, gfeat);
  }
  FRCNN_LAUNCH_CHECK(h, "roi_bwd_kernel");
  return FRCNN_OK;
}

}  // namespace frcnn
