/*
 * oracle/mscnn_oracle.c -- CPU restatement ("port") of the MS-CNN forward-path algorithms whose
 * inner loops are too slow in numpy.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * Every function names the reference file:line it restates.  Pinning: tests/test_oracle.py
 * checks each one against oracle/_ref (the reference's own code compiled verbatim) on random
 * and edge-case inputs, and against the golden vectors under tests/golden/ that were generated
 * from oracle/_ref (tests/golden/make_golden.py).  The reference itself ships no test for
 * BoxOutput / ROIPooling(pad_ratio) / BoxIOU.
 *
 * Compile: gcc -O2 -fPIC -shared -ffp-contract=off (no FMA contraction: the reference is built
 * -O2 for baseline x86-64, Makefile:318-322).  All arithmetic is fp32 unless noted.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- BoxIOU: src/caffe/util/math_functions.cpp:13-35.  mode 0 IOU, 1 IOMU, 2 IOFU. ---- */
static float box_iou(float x1, float y1, float w1, float h1, float x2, float y2, float w2, float h2,
                     int mode) {
  if (w1 <= 0 || h1 <= 0 || w2 <= 0 || h2 <= 0) return 0.0f;
  float tlx = x1 > x2 ? x1 : x2; /* std::max */
  float tly = y1 > y2 ? y1 : y2;
  float ex1 = x1 + w1, ex2 = x2 + w2, ey1 = y1 + h1, ey2 = y2 + h2;
  float brx = ex1 < ex2 ? ex1 : ex2; /* std::min */
  float bry = ey1 < ey2 ? ey1 : ey2;
  float over = (tlx >= brx || tly >= bry) ? 0.0f : (brx - tlx) * (bry - tly);
  float u;
  if (mode == 1) {
    float a1 = w1 * h1, a2 = w2 * h2;
    u = a1 < a2 ? a1 : a2;
  } else if (mode == 2) {
    u = w1 * h1;
  } else {
    u = w1 * h1 + w2 * h2 - over;
  }
  return over / u;
}

float oracle_box_iou(float x1, float y1, float w1, float h1, float x2, float y2, float w2, float h2,
                     int mode) {
  return box_iou(x1, y1, w1, h1, x2, y2, w2, h2, mode);
}

typedef struct {
  float score;
  int idx;
} score_idx;

/* std::greater<std::pair<Dtype,int>>: descending by score, then descending by index
 * (box_output_layer.cpp:168). */
static int cmp_score_idx_desc(const void* a, const void* b) {
  const score_idx* p = (const score_idx*)a;
  const score_idx* q = (const score_idx*)b;
  if (p->score > q->score) return -1;
  if (p->score < q->score) return 1;
  if (p->idx > q->idx) return -1;
  if (p->idx < q->idx) return 1;
  return 0;
}

/*
 * BoxOutputLayer::Forward_cpu: src/caffe/layers/box_output_layer.cpp:66-234, nmsMax :38-63.
 *   maps[j]: fp32 [num][channels][height[j]][width[j]]
 *   rois [cap][5], rois_score [cap][6] (may be NULL); per_image[num] receives the per-image
 *   proposal counts.  Returns the number of output rows (>= 1; 1 dummy row when empty), and
 *   *true_count the number of real proposals.
 */
int oracle_box_output(int num, int channels, int num_scales, const float* const* maps, const int* height,
                      const int* width, const unsigned* field_w, const unsigned* field_h,
                      const unsigned* downsample_rate, float fg_thr, float iou_thr, int nms_mode,
                      float field_whr, float field_xyr, float min_size, int max_nms_num,
                      int max_post_nms_num, int do_bbox_norm, const float* bbox_mean,
                      const float* bbox_std, float* rois, float* rois_score, int cap, int* per_image,
                      int* true_count) {
  const int cls_num = channels - 4;
  const float min_whr = logf(1.0f / field_whr), max_whr = logf(field_whr); /* :76 */
  const float min_xyr = -1.0f / field_xyr, max_xyr = 1.0f / field_xyr;      /* :77 */
  long total_anchors = 0;
  for (int j = 0; j < num_scales; ++j) total_anchors += (long)height[j] * width[j];
  float* boxes = (float*)malloc(sizeof(float) * 6 * (size_t)total_anchors);
  score_idx* order = (score_idx*)malloc(sizeof(score_idx) * (size_t)total_anchors);
  float* sorted = (float*)malloc(sizeof(float) * 6 * (size_t)total_anchors);
  char* keep = (char*)malloc((size_t)total_anchors);
  int out_rows = 0;

  for (int i = 0; i < num; ++i) { /* :107 */
    int bb_count = 0;
    for (int j = 0; j < num_scales; ++j) {
      const float fw = (float)field_w[j], fh = (float)field_h[j], rate = (float)downsample_rate[j];
      const int W = width[j], H = height[j];
      const int img_w = (int)(W * rate), img_h = (int)(H * rate); /* :115 int = int*float */
      const int spatial = W * H;
      const float* data = maps[j] + (size_t)i * channels * spatial;
      for (int id = 0; id < spatial; ++id) { /* :118 */
        const int h = id / W, w = id % W;
        float fg = -FLT_MAX;
        for (int k = 1; k < cls_num; ++k) {
          const float v = data[(size_t)k * spatial + id];
          fg = fg < v ? v : fg;
        }
        fg -= data[id]; /* :127 */
        if (!(fg >= fg_thr)) continue;
        const float* c = data + (size_t)cls_num * spatial + id;
        float bx = c[0], by = c[spatial], bw = c[2 * (size_t)spatial], bh = c[3 * (size_t)spatial];
        if (do_bbox_norm) { /* :138-143 */
          bx *= bbox_std[0]; by *= bbox_std[1]; bw *= bbox_std[2]; bh *= bbox_std[3];
          bx += bbox_mean[0]; by += bbox_mean[1]; bw += bbox_mean[2]; bh += bbox_mean[3];
        }
        bx = min_xyr < bx ? bx : min_xyr; bx = bx < max_xyr ? bx : max_xyr; /* :145 */
        by = min_xyr < by ? by : min_xyr; by = by < max_xyr ? by : max_xyr;
        bx = bx * fw + (w + 0.5f) * rate;
        by = by * fh + (h + 0.5f) * rate;
        bw = min_whr < bw ? bw : min_whr; bw = bw < max_whr ? bw : max_whr; /* :150 */
        bh = min_whr < bh ? bh : min_whr; bh = bh < max_whr ? bh : max_whr;
        bw = fw * expf(bw); bh = fh * expf(bh);
        bx = bx - bw / 2.0f; by = by - bh / 2.0f;
        bx = bx < 0.0f ? 0.0f : bx; by = by < 0.0f ? 0.0f : by; /* :154 */
        { float r = img_w - bx; bw = r < bw ? r : bw; }
        { float r = img_h - by; bh = r < bh ? r : bh; }
        if (bw >= min_size && bh >= min_size) { /* :157 */
          float* b = boxes + 6 * (size_t)bb_count;
          b[0] = (float)i; b[1] = bx; b[2] = by; b[3] = bw; b[4] = bh; b[5] = fg;
          order[bb_count].score = fg;
          order[bb_count].idx = bb_count;
          ++bb_count;
        }
      }
    }
    per_image[i] = 0;
    if (bb_count <= 0) continue; /* :166 */
    qsort(order, (size_t)bb_count, sizeof(score_idx), cmp_score_idx_desc); /* :168 */
    int n = bb_count;
    if (max_nms_num > 0 && bb_count > max_nms_num) n = max_nms_num; /* :176 */
    for (int k = 0; k < n; ++k) memcpy(sorted + 6 * (size_t)k, boxes + 6 * (size_t)order[k].idx, 6 * sizeof(float));
    /* nmsMax(greedy=true), :38-63 */
    for (int k = 0; k < n; ++k) keep[k] = 1;
    for (int a = 0; a < n; ++a) {
      if (!keep[a]) continue;
      const float* p = sorted + 6 * (size_t)a;
      for (int b = a + 1; b < n; ++b) {
        if (!keep[b]) continue;
        const float* q = sorted + 6 * (size_t)b;
        if (box_iou(p[1], p[2], p[3], p[4], q[1], q[2], q[3], q[4], nms_mode) > iou_thr) keep[b] = 0;
      }
    }
    int kept = 0;
    for (int k = 0; k < n; ++k) {
      if (!keep[k]) continue;
      if (max_post_nms_num > 0 && kept >= max_post_nms_num) break; /* :184-186 */
      if (out_rows < cap) {
        const float* p = sorted + 6 * (size_t)k;
        float* r = rois + 5 * (size_t)out_rows; /* :201-211 */
        r[0] = p[0]; r[1] = p[1]; r[2] = p[2]; r[3] = p[1] + p[3]; r[4] = p[2] + p[4];
        if (rois_score) {
          float* s = rois_score + 6 * (size_t)out_rows;
          s[0] = p[0]; s[1] = p[1]; s[2] = p[2]; s[3] = p[1] + p[3]; s[4] = p[2] + p[4]; s[5] = p[5];
        }
      }
      ++out_rows;
      ++kept;
    }
    per_image[i] = kept;
  }
  free(boxes); free(order); free(sorted); free(keep);
  *true_count = out_rows;
  if (out_rows <= 0) { /* :195-199, :214-218 */
    rois[0] = 0; rois[1] = 1; rois[2] = 1; rois[3] = 10; rois[4] = 10;
    if (rois_score) memset(rois_score, 0, 6 * sizeof(float));
    return 1;
  }
  return out_rows;
}

/*
 * ROIPoolingLayer::Forward_cpu: src/caffe/layers/roi_pooling_layer.cpp:49-139.
 *   data [batch][channels][height][width], rois [num_rois][5], top [num_rois][channels][ph][pw].
 */
void oracle_roi_pool(const float* data, int batch, int channels, int height, int width, const float* rois,
                     int num_rois, int pooled_h, int pooled_w, float spatial_scale, float pad_ratio,
                     float* top) {
  for (int n = 0; n < num_rois; ++n) {
    const float* r = rois + 5 * (size_t)n;
    int b = (int)r[0];
    if (b < 0) b = 0;
    if (b >= batch) b = batch - 1;
    const float pad_w = (r[3] - r[1] + 1) * pad_ratio; /* :69-70 */
    const float pad_h = (r[4] - r[2] + 1) * pad_ratio;
    const int sw = (int)roundf((r[1] - pad_w) * spatial_scale);
    const int sh = (int)roundf((r[2] - pad_h) * spatial_scale);
    const int ew = (int)roundf((r[3] + pad_w) * spatial_scale);
    const int eh = (int)roundf((r[4] + pad_h) * spatial_scale);
    int roi_h = eh - sh + 1; if (roi_h < 1) roi_h = 1;
    int roi_w = ew - sw + 1; if (roi_w < 1) roi_w = 1;
    const float bin_h = (float)roi_h / (float)pooled_h;
    const float bin_w = (float)roi_w / (float)pooled_w;
    for (int c = 0; c < channels; ++c) {
      const float* plane = data + ((size_t)b * channels + c) * height * width;
      float* out = top + ((size_t)n * channels + c) * pooled_h * pooled_w;
      for (int ph = 0; ph < pooled_h; ++ph)
        for (int pw = 0; pw < pooled_w; ++pw) {
          int hs = (int)floorf((float)ph * bin_h), ws = (int)floorf((float)pw * bin_w);
          int he = (int)ceilf((float)(ph + 1) * bin_h), we = (int)ceilf((float)(pw + 1) * bin_w);
          hs += sh; he += sh; ws += sw; we += sw;
          hs = hs < 0 ? 0 : (hs > height ? height : hs);
          he = he < 0 ? 0 : (he > height ? height : he);
          ws = ws < 0 ? 0 : (ws > width ? width : ws);
          we = we < 0 ? 0 : (we > width ? width : we);
          float best = (he <= hs || we <= ws) ? 0.0f : -FLT_MAX;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
              const float v = plane[(size_t)h * width + w];
              if (v > best) best = v;
            }
          out[ph * pooled_w + pw] = best;
        }
    }
  }
}

/*
 * PoolingLayer::Forward_cpu, MAX (:128-187) and AVE (:188-220) with the ceil-mode output size of
 * Reshape (:79-123): src/caffe/layers/pooling_layer.cpp.  mode 0 = MAX, 1 = AVE.
 */
void oracle_pool(const float* x, int num, int channels, int height, int width, int kernel, int stride,
                 int pad, int mode, float* y, int* out_h, int* out_w) {
  int ph = (int)ceilf((float)(height + 2 * pad - kernel) / stride) + 1;
  int pw = (int)ceilf((float)(width + 2 * pad - kernel) / stride) + 1;
  if (pad) { /* :94-103 */
    if ((ph - 1) * stride >= height + pad) --ph;
    if ((pw - 1) * stride >= width + pad) --pw;
  }
  *out_h = ph; *out_w = pw;
  if (!y) return;
  for (size_t nc = 0; nc < (size_t)num * channels; ++nc) {
    const float* in = x + nc * height * width;
    float* out = y + nc * ph * pw;
    for (int oh = 0; oh < ph; ++oh)
      for (int ow = 0; ow < pw; ++ow) {
        int hs = oh * stride - pad, ws = ow * stride - pad;
        if (mode == 0) {
          int he = hs + kernel < height ? hs + kernel : height;
          int we = ws + kernel < width ? ws + kernel : width;
          if (hs < 0) hs = 0;
          if (ws < 0) ws = 0;
          float best = -FLT_MAX;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w)
              if (in[h * width + w] > best) best = in[h * width + w];
          out[oh * pw + ow] = best;
        } else {
          int he = hs + kernel < height + pad ? hs + kernel : height + pad; /* :196-203 */
          int we = ws + kernel < width + pad ? ws + kernel : width + pad;
          const int pool_size = (he - hs) * (we - ws);
          if (hs < 0) hs = 0;
          if (ws < 0) ws = 0;
          if (he > height) he = height;
          if (we > width) we = width;
          float acc = 0;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) acc += in[h * width + w];
          out[oh * pw + ow] = acc / pool_size;
        }
      }
  }
}

/*
 * DeconvolutionLayer::Forward_cpu for group == channels (depthwise), no bias:
 * src/caffe/layers/deconv_layer.cpp:25-40 = backward_cpu_gemm (col = w^T x) + col2im
 * (util/im2col.cpp col2im_cpu: accumulate kernel offsets in (ky,kx) order into a zeroed image).
 *   x [num][channels][h][w], wt [channels][1][k][k], y [num][channels][ho][wo],
 *   ho = stride*(h-1) + k - 2*pad (deconv_layer.cpp:8-22).
 */
void oracle_deconv_depthwise(const float* x, int num, int channels, int height, int width,
                             const float* wt, int kernel, int stride, int pad, float* y) {
  const int ho = stride * (height - 1) + kernel - 2 * pad;
  const int wo = stride * (width - 1) + kernel - 2 * pad;
  for (size_t nc = 0; nc < (size_t)num * channels; ++nc) {
    const int c = (int)(nc % channels);
    const float* in = x + nc * height * width;
    float* out = y + nc * ho * wo;
    memset(out, 0, sizeof(float) * ho * wo);
    for (int ky = 0; ky < kernel; ++ky)
      for (int kx = 0; kx < kernel; ++kx) {
        const float wv = wt[((size_t)c * kernel + ky) * kernel + kx];
        for (int iy = 0; iy < height; ++iy) {
          const int oy = iy * stride - pad + ky;
          if (oy < 0 || oy >= ho) continue;
          for (int ix = 0; ix < width; ++ix) {
            const int ox = ix * stride - pad + kx;
            if (ox < 0 || ox >= wo) continue;
            out[oy * wo + ox] += wv * in[iy * width + ix];
          }
        }
      }
  }
}

/*
 * bbNms 'maxg' with 'union' denominator: utils/bbNms.m:112-126 (nmsMax, greedy), in double like
 * MATLAB.  bbs [n][5] = [x y w h score]; `order` receives the kept row indices (into bbs) in
 * descending-score order; returns their count.  MATLAB's sort(...,'descend') is stable.
 */
typedef struct {
  double score;
  int idx;
} dscore_idx;
static int cmp_dscore_desc_stable(const void* a, const void* b) {
  const dscore_idx* p = (const dscore_idx*)a;
  const dscore_idx* q = (const dscore_idx*)b;
  if (p->score > q->score) return -1;
  if (p->score < q->score) return 1;
  return (p->idx > q->idx) - (p->idx < q->idx);
}
int oracle_bbnms_maxg(const double* bbs, int n, double overlap, int* order) {
  dscore_idx* ord = (dscore_idx*)malloc(sizeof(dscore_idx) * (size_t)(n > 0 ? n : 1));
  char* kp = (char*)malloc((size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) { ord[i].score = bbs[5 * (size_t)i + 4]; ord[i].idx = i; kp[i] = 1; }
  qsort(ord, (size_t)n, sizeof(dscore_idx), cmp_dscore_desc_stable);
  for (int i = 0; i < n; ++i) {
    if (!kp[i]) continue;
    const double* a = bbs + 5 * (size_t)ord[i].idx;
    const double axe = a[0] + a[2], aye = a[1] + a[3], aa = a[2] * a[3];
    for (int j = i + 1; j < n; ++j) {
      if (!kp[j]) continue;
      const double* b = bbs + 5 * (size_t)ord[j].idx;
      const double bxe = b[0] + b[2], bye = b[1] + b[3];
      const double iw = (axe < bxe ? axe : bxe) - (a[0] > b[0] ? a[0] : b[0]);
      if (iw <= 0) continue;
      const double ih = (aye < bye ? aye : bye) - (a[1] > b[1] ? a[1] : b[1]);
      if (ih <= 0) continue;
      double o = iw * ih;
      const double u = aa + b[2] * b[3] - o;
      o = o / u;
      if (o > overlap) kp[j] = 0;
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (kp[i]) order[m++] = ord[i].idx;
  free(ord); free(kp);
  return m;
}

/*
 * ROIAlignLayer::Forward_cpu: src/caffe/layers/roi_align_layer.cpp:49-139.  Bilinear samples on
 * the (ph+1) x (pw+1) grid of bin corners of the padded ROI; malformed ROI or a grid point outside
 * the map gives 0.  y [rois][channels][ph+1][pw+1].
 */
void oracle_roi_align(const float* data, int batch, int channels, int height, int width, const float* rois,
                      int num_rois, int pooled_h, int pooled_w, float spatial_scale, float pad_ratio, float* y) {
  const int gh = pooled_h + 1, gw = pooled_w + 1;
  for (int n = 0; n < num_rois; ++n) {
    const float* r = rois + (size_t)n * 5;
    int b = (int)r[0];
    if (b < 0) b = 0;
    if (b >= batch) b = batch - 1; /* CHECKed at :63-64 */
    float pad_w = (r[3] - r[1] + 1) * pad_ratio; /* :67-68 */
    float pad_h = (r[4] - r[2] + 1) * pad_ratio;
    float start_w = (r[1] - pad_w) * spatial_scale; /* :71-74 */
    float start_h = (r[2] - pad_h) * spatial_scale;
    float end_w = (r[3] + pad_w) * spatial_scale;
    float end_h = (r[4] + pad_h) * spatial_scale;
    start_w -= 0.5; start_h -= 0.5; end_w -= 0.5; end_h -= 0.5; /* :77-78 */
    const float roi_h = end_h - start_h, roi_w = end_w - start_w;
    const float bin_h = roi_h / (float)pooled_h, bin_w = roi_w / (float)pooled_w;
    for (int c = 0; c < channels; ++c) {
      const float* in = data + ((size_t)b * channels + c) * height * width;
      float* out = y + ((size_t)n * channels + c) * gh * gw;
      for (int ph = 0; ph < gh; ++ph)
        for (int pw = 0; pw < gw; ++pw) {
          float* o = out + ph * gw + pw;
          if (roi_h <= 0 || roi_w <= 0) { *o = 0; continue; } /* :94-97 */
          float hf = start_h + (float)ph * bin_h; /* :100-101 */
          float wf = start_w + (float)pw * bin_w;
          if (hf < -0.5 || hf > (height - 0.5) || wf < -0.5 || wf > (width - 0.5)) { *o = 0; continue; }
          int h0 = (int)floorf(hf), w0 = (int)floorf(wf); /* :111-112 */
          int h1 = h0 + 1, w1 = w0 + 1;
          hf = fminf(fmaxf(hf, 0.0f), (float)(height - 1)); /* :115-120 */
          wf = fminf(fmaxf(wf, 0.0f), (float)(width - 1));
          h0 = h0 < 0 ? 0 : (h0 > height - 1 ? height - 1 : h0);
          w0 = w0 < 0 ? 0 : (w0 > width - 1 ? width - 1 : w0);
          h1 = h1 < 0 ? 0 : (h1 > height - 1 ? height - 1 : h1);
          w1 = w1 < 0 ? 0 : (w1 > width - 1 ? width - 1 : w1);
          const float lh = hf - h0, lw = wf - w0; /* :123-124 */
          const float hh = 1 - lh, hw = 1 - lw;
          const float w00 = hw * hh, w10 = lw * hh, w01 = hw * lh, w11 = lw * lh;
          const float v00 = in[h0 * width + w0], v10 = in[h0 * width + w1];
          const float v01 = in[h1 * width + w0], v11 = in[h1 * width + w1];
          *o = w00 * v00 + w10 * v10 + w01 * v01 + w11 * v11; /* :137 */
        }
    }
  }
}

/*
 * DecodeBBoxLayer::Forward_cpu, TEST phase (src/caffe/layers/decode_bbox_layer.cpp:53-124) with
 * DecodeBBoxesWithPrior (src/caffe/util/math_functions.cpp:46-77): row i of `out` is
 * [img, x1, y1, x2, y2] of prior i moved by the class-1 deltas bbox[i][4..8).
 */
void oracle_decode_bbox(const float* bbox, const float* prior, int num, int bbox_dim, const float* mean,
                        const float* stdv, float* out) {
  for (int i = 0; i < num; ++i) {
    const float* q = prior + (size_t)i * 5;
    const float xmin = q[1], ymin = q[2], xmax = q[3], ymax = q[4];
    const float pw = xmax - xmin + 1, ph = ymax - ymin + 1; /* math_functions.cpp:54-55 */
    const float cx = 0.5 * (xmax + xmin), cy = 0.5 * (ymax + ymin);
    const float* d = bbox + (size_t)i * bbox_dim + 4; /* class 1: decode_bbox_layer.cpp:115 */
    const float bx = d[0] * stdv[0] + mean[0], by = d[1] * stdv[1] + mean[1];
    const float bw = d[2] * stdv[2] + mean[2], bh = d[3] * stdv[3] + mean[3];
    float tx = bx * pw + cx, ty = by * ph + cy; /* :67-68 */
    const float tw = pw * expf(bw), th = ph * expf(bh);
    tx -= (tw - 1) / 2; ty -= (th - 1) / 2;
    float* o = out + (size_t)i * 5;
    o[0] = q[0];
    o[1] = tx; o[2] = ty;
    o[3] = tx + tw - 1; o[4] = ty + th - 1; /* :72-73 */
  }
}
