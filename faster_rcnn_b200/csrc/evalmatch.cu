// VOC detection evaluation on the device (widening row): replaces the matching loop and the
// precision / recall / 11-point AP arithmetic of eval_dets.voc_eval (eval_dets.py:75-125, :8-19).
//
//  voc_match_kernel   one warp per image.  The image's detections are visited in descending-confidence order
//                     (their ranks in the globally sorted list come in as a CSR); lanes split the image's GT
//                     boxes, IoU in float64 with the devkit's +1 convention (eval_dets.py:89-102), a warp
//                     arg-max with first-index ties (np.argmax), then the greedy "already detected" marking
//                     (eval_dets.py:107-116).  Images are independent, the order inside an image is sequential
//                     by definition.
//  voc_pr_ap_kernel   one CTA: chunked inclusive scans of tp / fp (np.cumsum), rec = tp/npos,
//                     prec = tp/max(tp+fp, eps), then for every threshold the masked maximum
//                     max(prec[rec >= t]) and ap = sum(p/11) in threshold order (float64, like numpy).
#include "common.cuh"

namespace frcnn {

__global__ void __launch_bounds__(256)
voc_match_kernel(const double* __restrict__ det_boxes, const int* __restrict__ img_det_offsets,
                 const int* __restrict__ img_det_rank, const double* __restrict__ gt_boxes,
                 const unsigned char* __restrict__ gt_difficult, const int* __restrict__ img_gt_offsets, int n_img,
                 double ovthresh, unsigned char* __restrict__ gt_taken, double* __restrict__ tp,
                 double* __restrict__ fp) {
  const int lane = threadIdx.x & 31;
  const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (img >= n_img) return;                               // warp-uniform
  const int d0 = img_det_offsets[img], d1 = img_det_offsets[img + 1];
  const int g0 = img_gt_offsets[img], g1 = img_gt_offsets[img + 1];
  for (int j = g0 + lane; j < g1; j += 32) gt_taken[j] = 0;
  __syncwarp();
  for (int d = d0; d < d1; ++d) {
    const int rank = img_det_rank[d];
    const double bx1 = det_boxes[4 * (size_t)rank], by1 = det_boxes[4 * (size_t)rank + 1];
    const double bx2 = det_boxes[4 * (size_t)rank + 2], by2 = det_boxes[4 * (size_t)rank + 3];
    const double b_area = __dmul_rn(__dadd_rn(__dsub_rn(bx2, bx1), 1.0), __dadd_rn(__dsub_rn(by2, by1), 1.0));
    double best = -INFINITY;
    int best_j = 0x7fffffff;
    for (int j = g0 + lane; j < g1; j += 32) {
      const double gx1 = gt_boxes[4 * (size_t)j], gy1 = gt_boxes[4 * (size_t)j + 1];
      const double gx2 = gt_boxes[4 * (size_t)j + 2], gy2 = gt_boxes[4 * (size_t)j + 3];
      const double iw = fmax(__dadd_rn(__dsub_rn(fmin(gx2, bx2), fmax(gx1, bx1)), 1.0), 0.0);
      const double ih = fmax(__dadd_rn(__dsub_rn(fmin(gy2, by2), fmax(gy1, by1)), 1.0), 0.0);
      const double inters = __dmul_rn(iw, ih);
      const double g_area = __dmul_rn(__dadd_rn(__dsub_rn(gx2, gx1), 1.0), __dadd_rn(__dsub_rn(gy2, gy1), 1.0));
      const double uni = __dsub_rn(__dadd_rn(b_area, g_area), inters);
      const double ov = __ddiv_rn(inters, uni);
      if (ov > best) { best = ov; best_j = j; }             // ascending j: the first maximum stays
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {                      // (max overlap, min index) across the warp
      const double ob = __shfl_xor_sync(0xffffffffu, best, s);
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, s);
      if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
    }
    if (lane == 0) {
      double t = 0.0, f = 0.0;
      if (g1 > g0 && best > ovthresh) {
        if (!gt_difficult[best_j]) {
          if (!gt_taken[best_j]) { t = 1.0; gt_taken[best_j] = 1; }
          else f = 1.0;
        }
      } else {
        f = 1.0;
      }
      tp[rank] = t;
      fp[rank] = f;
    }
    __syncwarp();
  }
}

constexpr int PR_THREADS = 1024;

__global__ void __launch_bounds__(PR_THREADS, 1)
voc_pr_ap_kernel(const double* __restrict__ tp, const double* __restrict__ fp, int nd, double npos,
                 const double* __restrict__ thresholds, int n_thr, double* __restrict__ rec,
                 double* __restrict__ prec, double* __restrict__ ap) {
  __shared__ double s_wt[32], s_wf[32];
  __shared__ double s_carry_t, s_carry_f;
  __shared__ double s_max[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_carry_t = 0.0; s_carry_f = 0.0; }
  __syncthreads();
  // np.cumsum: counts are small integers, so the sums are exact in float64 in any order
  for (int base = 0; base < nd; base += PR_THREADS) {
    const int i = base + tid;
    double t = (i < nd) ? tp[i] : 0.0, f = (i < nd) ? fp[i] : 0.0;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const double ut = __shfl_up_sync(0xffffffffu, t, s), uf = __shfl_up_sync(0xffffffffu, f, s);
      if (lane >= s) { t += ut; f += uf; }
    }
    if (lane == 31) { s_wt[warp] = t; s_wf[warp] = f; }
    __syncthreads();
    double bt = s_carry_t, bf = s_carry_f;
    for (int w = 0; w < warp; ++w) { bt += s_wt[w]; bf += s_wf[w]; }
    t += bt;
    f += bf;
    if (i < nd) {
      rec[i] = __ddiv_rn(t, npos);
      prec[i] = __ddiv_rn(t, fmax(__dadd_rn(t, f), 2.220446049250313e-16));      // np.finfo(np.float64).eps
    }
    __syncthreads();
    if (tid == PR_THREADS - 1) { s_carry_t = t; s_carry_f = f; }
    __syncthreads();
  }
  // 11-point AP (eval_dets.py:10-17): p = max(prec[rec >= t]) or 0 when empty; ap += p / 11 in threshold order
  double total = 0.0;
  for (int k = 0; k < n_thr; ++k) {
    const double thr = thresholds[k];
    double m = -1.0;                                      // precision is >= 0: -1 marks "no element"
    for (int i = tid; i < nd; i += PR_THREADS)
      if (rec[i] >= thr) m = fmax(m, prec[i]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    if (tid == 0) {
      double mm = -1.0;
      for (int w = 0; w < PR_THREADS / 32; ++w) mm = fmax(mm, s_max[w]);
      total = __dadd_rn(total, __ddiv_rn(mm < 0.0 ? 0.0 : mm, (double)n_thr));
    }
    __syncthreads();
  }
  if (tid == 0) *ap = total;
}

int launch_voc_match(frcnn_handle* h, cudaStream_t stream, const double* det_boxes, const int32_t* img_det_offsets,
                     const int32_t* img_det_rank, const double* gt_boxes, const uint8_t* gt_difficult,
                     const int32_t* img_gt_offsets, int n_img, double ovthresh, uint8_t* gt_taken, double* tp,
                     double* fp) {
  voc_match_kernel<<<(n_img + 7) / 8, 256, 0, stream>>>(det_boxes, img_det_offsets, img_det_rank, gt_boxes, gt_difficult,
                                                       img_gt_offsets, n_img, ovthresh, gt_taken, tp, fp);
  FRCNN_LAUNCH_CHECK(h, "voc_match_kernel");
  return FRCNN_OK;
}

int launch_voc_pr_ap(frcnn_handle* h, cudaStream_t stream, const double* tp, const double* fp, int nd, double npos,
                     const double* thresholds, int n_thr, double* rec, double* prec, double* ap) {
  voc_pr_ap_kernel<<<1, PR_THREADS, 0, stream>>>(tp, fp, nd, npos, thresholds, n_thr, rec, prec, ap);
  FRCNN_LAUNCH_CHECK(h, "voc_pr_ap_kernel");
  return FRCNN_OK;
}

}  // namespace frcnn
