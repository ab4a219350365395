/*
 * fe_oracle.c — CPU restatement of the per-frame front end.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Pinned against src/PatchFinder.cc / MiniPatch.cc / ShiTomasi.cc compiled from the reference (tests/test_oracle_vs_ref.py); libCVD (halfSample, FAST, transform) restated [3P] and pinned against OpenCV.
 *
 * Follows (file:line in /root/reference):
 *   src/KeyFrame.cc:189-190     CVD::halfSample                       [3P libCVD]  -> ora_halfsample
 *   src/KeyFrame.cc:259-262     fast_corner_detect_10 / _score_10      [3P libCVD]  -> ora_fast10_*
 *   src/KeyFrame.cc:264-315     histogram, adaptive threshold, mask filter         -> ora_level_corners
 *   src/KeyFrame.cc:348-355     row LUT                                            -> ora_level_corners
 *   src/ShiTomasi.cc:34-63      FindShiTomasiScoreAtPoint                          -> ora_shitomasi
 *   src/PatchFinder.cc:135-182  MakeTemplateCoarseCont -> CVD::transform [3P]      -> ora_patch_template
 *   src/PatchFinder.cc:209-223  MakeTemplateSums
 *   src/PatchFinder.cc:229-355  FindPatchCoarse                                    -> ora_find_patch_coarse
 *   src/PatchFinder.cc:362-470  MakeSubPixTemplate / IterateSubPix(ToConvergence)  -> ora_subpix
 *   src/PatchFinder.cc:511-658  ZMSSDAtPoint                                       -> ora_zmssd
 *   src/MiniPatch.cc:34-113     SSDAtPoint / FindPatch                             -> ora_minipatch_*
 * [3P] libCVD semantics restated from libCVD release 20121025 (vision.h halfSample/transform/sample,
 *      fast_corner.h): ring layout, strict > / < comparisons, 3 px border, raster order, truncating
 *      2x2 mean, truncating float->byte conversion in sample().
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* [3P] CVD::halfSample generic template: truncating mean of the 2x2 block, out size = in size / 2 */
void ora_halfsample(const uint8_t* in, int w, int h, int in_stride, uint8_t* out, int out_stride)
{
  const int ow = w / 2, oh = h / 2;
  for (int y = 0; y < oh; y++) {
    const uint8_t* r0 = in + (size_t)(2 * y) * in_stride;
    const uint8_t* r1 = r0 + in_stride;
    uint8_t* o = out + (size_t)y * out_stride;
    for (int x = 0; x < ow; x++) o[x] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1]) / 4);
  }
}

/* [3P] FAST ring: radius-3 Bresenham circle, 16 pixels, clockwise from 12 o'clock
   (image y grows downward) */
static const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int RING_DY[16] = { -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3 };

static int is_cornerN(const uint8_t* p, int stride, int b, int n_arc)
{
  const int c = *p;
  for (int pol = 0; pol < 2; pol++)
    for (int s = 0; s < 16; s++) {
      int k;
      for (k = 0; k < n_arc; k++) {
        const int i = (s + k) & 15;
        const int v = p[RING_DY[i] * stride + RING_DX[i]];
        if (pol == 0 ? !(v > c + b) : !(v < c - b)) break;
      }
      if (k == n_arc) return 1;
    }
  return 0;
}
int ora_fastN_detect_bruteforce(const uint8_t* im, int w, int h, int stride, int b, int n_arc, int32_t* xy, int cap)
{
  int n = 0;
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++)
      if (is_cornerN(im + (size_t)y * stride + x, stride, b, n_arc)) {
        if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; }
        n++;
      }
  return n;
}
int ora_fast10_detect_bruteforce(const uint8_t* im, int w, int h, int stride, int b, int32_t* xy, int cap)
{
  return ora_fastN_detect_bruteforce(im, w, h, stride, b, 10, xy, cap);
}

/* contiguous run of >= n set bits in a 16-bit circular mask */
static int has_run(unsigned m, int n)
{
  unsigned d = m | (m << 16);
  unsigned r = d;
  for (int k = 1; k < n; k++) r &= (d >> k);
  return (r & 0xFFFFu) != 0;
}
/* Same definition as the brute force, with the usual early rejection on the compass points. */
int ora_fast10_detect(const uint8_t* im, int w, int h, int stride, int b, int32_t* xy, int cap)
{
  int n = 0;
  int off[16];
  for (int i = 0; i < 16; i++) off[i] = RING_DY[i] * stride + RING_DX[i];
  for (int y = 3; y < h - 3; y++) {
    const uint8_t* row = im + (size_t)y * stride;
    for (int x = 3; x < w - 3; x++) {
      const uint8_t* p = row + x;
      const int cb = *p + b, c_b = *p - b;
      /* any 10-arc contains one of {0,8} and one of {4,12} */
      const int v0 = p[off[0]], v8 = p[off[8]];
      if (!(v0 > cb || v0 < c_b || v8 > cb || v8 < c_b)) continue;
      const int v4 = p[off[4]], v12 = p[off[12]];
      if (!(v4 > cb || v4 < c_b || v12 > cb || v12 < c_b)) continue;
      unsigned br = 0, dk = 0;
      for (int i = 0; i < 16; i++) {
        const int v = p[off[i]];
        br |= (unsigned)(v > cb) << i;
        dk |= (unsigned)(v < c_b) << i;
      }
      if (has_run(br, 10) || has_run(dk, 10)) {
        if (n < cap) { xy[2 * n] = x; xy[2 * n + 1] = y; }
        n++;
      }
    }
  }
  return n;
}

/* [3P] fast_corner_score_10: largest threshold t >= b for which the pixel is still a FAST-10 corner.
   Closed form: max over arcs of (min over arc of signed difference) - 1. */
void ora_fast10_score(const uint8_t* im, int stride, const int32_t* xy, int n, int b, int32_t* scores)
{
  for (int c = 0; c < n; c++) {
    const uint8_t* p = im + (size_t)xy[2 * c + 1] * stride + xy[2 * c];
    const int ctr = *p;
    int d[16];
    for (int i = 0; i < 16; i++) d[i] = p[RING_DY[i] * stride + RING_DX[i]] - ctr;
    int best = b;
    for (int s = 0; s < 16; s++) {
      int mn = 255, mx = -255;
      for (int k = 0; k < 10; k++) {
        const int v = d[(s + k) & 15];
        if (v < mn) mn = v;
        if (v > mx) mx = v;
      }
      if (mn - 1 > best) best = mn - 1;       /* all brighter than ctr + t  <=>  t < min diff */
      if (-mx - 1 > best) best = -mx - 1;     /* all darker  than ctr - t  <=>  t < min(-diff) */
    }
    scores[c] = best;
  }
}
/* The bisection exactly as libCVD performs it (bmin = b, bmax = 255). */
void ora_fast10_score_bisect(const uint8_t* im, int stride, const int32_t* xy, int n, int b, int32_t* scores)
{
  for (int c = 0; c < n; c++) {
    const uint8_t* p = im + (size_t)xy[2 * c + 1] * stride + xy[2 * c];
    int bmin = b, bmax = 255, t = (bmax + bmin) / 2;
    for (;;) {
      if (is_cornerN(p, stride, t, 10)) bmin = t; else bmax = t;
      if (bmin == bmax - 1 || bmin == bmax) break;
      t = (bmin + bmax) / 2;
    }
    scores[c] = bmin;
  }
}

/* src/KeyFrame.cc:247-355 for one level */
int ora_level_corners(const uint8_t* im, int w, int h, int stride, const uint8_t* mask, int mask_stride,
                      int adaptive, int fixed_thresh, int32_t* xy, int cap, int32_t* freq, int32_t* thresh_out,
                      int32_t* row_lut)
{
  int n_keep = 0;
  for (int t = 0; t <= 30; t++) freq[t] = 0;
  if (adaptive) {
    int n = ora_fast10_detect(im, w, h, stride, 5, NULL, 0);
    int32_t* all = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)(n > 0 ? n : 1));
    int32_t* sc = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    ora_fast10_detect(im, w, h, stride, 5, all, n);
    ora_fast10_score(im, stride, all, n, 5, sc);
    for (int j = 0; j < n; j++)                                   /* :264-275 */
      for (int t = 5; t <= 30; ++t) {
        if (sc[j] >= t) freq[t]++;
        if (sc[j] == t) break;
      }
    const double target = -1 * (w * h) / 500.0;                    /* :279 */
    int thr = 5;
    for (int t = 5; t <= 30; ++t) {                                /* :283-300 */
      double deriv;
      if (t == 5) deriv = freq[t + 1] - freq[t];
      else if (t == 30) deriv = freq[t] - freq[t - 1];
      else deriv = (freq[t + 1] - freq[t - 1]) / 2.0;
      thr = t;
      if (deriv > target) break;
    }
    for (int j = 0; j < n; j++) {                                  /* :303-312 */
      if (mask && mask[(size_t)all[2 * j + 1] * mask_stride + all[2 * j]] < 255) continue;
      if (sc[j] < thr) continue;
      if (n_keep < cap) { xy[2 * n_keep] = all[2 * j]; xy[2 * n_keep + 1] = all[2 * j + 1]; }
      n_keep++;
    }
    *thresh_out = thr;
    free(all); free(sc);
  } else {
    n_keep = ora_fast10_detect(im, w, h, stride, fixed_thresh, xy, cap);
    *thresh_out = fixed_thresh;
  }
  if (row_lut) {                                                   /* :348-355 */
    int v = 0;
    const int nk = n_keep < cap ? n_keep : cap;
    for (int y = 0; y < h; y++) {
      while (v < nk && y > xy[2 * v + 1]) v++;
      row_lut[y] = v;
    }
  }
  return n_keep;
}

/* src/ShiTomasi.cc:34-63 */
double ora_shitomasi(const uint8_t* im, int stride, int hb, int cx, int cy)
{
  double dXX = 0, dYY = 0, dXY = 0;
  for (int y = cy - hb; y <= cy + hb; y++)
    for (int x = cx - hb; x <= cx + hb; x++) {
      const double dx = im[(size_t)y * stride + x + 1] - im[(size_t)y * stride + x - 1];
      const double dy = im[(size_t)(y + 1) * stride + x] - im[(size_t)(y - 1) * stride + x];
      dXX += dx * dx; dYY += dy * dy; dXY += dx * dy;
    }
  const int nPixels = (2 * hb + 1) * (2 * hb + 1);
  dXX = dXX / (2.0 * nPixels);
  dYY = dYY / (2.0 * nPixels);
  dXY = dXY / (2.0 * nPixels);
  return 0.5 * (dXX + dYY - sqrt((dXX + dYY) * (dXX + dYY) - 4 * (dXX * dYY - dXY * dXY)));
}

/* [3P] CVD::sample<byte,byte>: bilinear in double, implicit double->byte truncation */
static uint8_t sample_u8(const uint8_t* im, int stride, double x, double y)
{
  const int lx = (int)x, ly = (int)y;
  x -= lx; y -= ly;
  const uint8_t* p = im + (size_t)ly * stride + lx;
  const double v = (1 - y) * ((1 - x) * p[0] + x * p[1]) + y * ((1 - x) * p[stride] + x * p[stride + 1]);
  return (uint8_t)v;
}
/* [3P] CVD::transform(in, out(8x8), M, inOrig, outOrig=(4,4)).  Returns number of pixels outside. */
int ora_patch_template(const uint8_t* src, int w, int h, int stride, const double* M, double cx, double cy,
                       uint8_t* out)
{
  const int ow = 8, oh = 8;
  const double across[2] = { M[0], M[2] }, down[2] = { M[1], M[3] };
  const double oo[2] = { 4, 4 };
  double p0[2] = { cx - (M[0] * oo[0] + M[1] * oo[1]), cy - (M[2] * oo[0] + M[3] * oo[1]) };
  double min_x = p0[0], min_y = p0[1], max_x = p0[0], max_y = p0[1];
  if (across[0] < 0) min_x += ow * across[0]; else max_x += ow * across[0];
  if (down[0] < 0) min_x += oh * down[0]; else max_x += oh * down[0];
  if (across[1] < 0) min_y += ow * across[1]; else max_y += ow * across[1];
  if (down[1] < 0) min_y += oh * down[1]; else max_y += oh * down[1];
  const double cr[2] = { down[0] - ow * across[0], down[1] - ow * across[1] };
  double p[2] = { p0[0], p0[1] };
  if (min_x >= 0 && min_y >= 0 && max_x < w - 1 && max_y < h - 1) {
    for (int i = 0; i < oh; ++i, p[0] += cr[0], p[1] += cr[1])
      for (int j = 0; j < ow; ++j, p[0] += across[0], p[1] += across[1]) out[i * 8 + j] = sample_u8(src, stride, p[0], p[1]);
    return 0;
  }
  const double xb = w - 1, yb = h - 1;
  int count = 0;
  for (int i = 0; i < oh; ++i, p[0] += cr[0], p[1] += cr[1])
    for (int j = 0; j < ow; ++j, p[0] += across[0], p[1] += across[1]) {
      if (0 <= p[0] && 0 <= p[1] && p[0] < xb && p[1] < yb) out[i * 8 + j] = sample_u8(src, stride, p[0], p[1]);
      else { out[i * 8 + j] = 0; ++count; }
    }
  return count;
}

/* src/PatchFinder.cc:511-658 (scalar branch; the SSE branch computes the same integers) */
int ora_zmssd(const uint8_t* im, int w, int h, int stride, const uint8_t* t, int tsum, int tsumsq, int x, int y,
              int max_ssd)
{
  if (!(x >= 4 && y >= 4 && x < w - 4 && y < h - 4)) return max_ssd + 1;   /* in_image_with_border(ir, 4) */
  int nImageSumSq = 0, nImageSum = 0, nCrossSum = 0;
  for (int r = 0; r < 8; r++) {
    const uint8_t* ip = im + (size_t)(y - 4 + r) * stride + (x - 4);
    const uint8_t* tp = t + r * 8;
    for (int c = 0; c < 8; c++) {
      const int n = ip[c];
      nImageSum += n; nImageSumSq += n * n; nCrossSum += n * tp[c];
    }
  }
  const int SA = tsum, SB = nImageSum, N = 64;
  return ((2 * SA * SB - SA * SA - SB * SB) / N + nImageSumSq + tsumsq - 2 * nCrossSum);
}

/* src/PatchFinder.cc:229-355.  corners in level coordinates, raster order, with row LUT. */
int ora_find_patch_coarse(const uint8_t* im, int w, int h, int stride, const int32_t* cxy, int nc,
                          const int32_t* lut, const uint8_t* t, int level, int px, int py, int range_in,
                          int exhaustive, int32_t* best_xy, int32_t* score)
{
  const int max_ssd = 8 * 8 * 250;                               /* :44,61 */
  int tsum = 0, tsumsq = 0;
  for (int i = 0; i < 64; i++) { tsum += t[i]; tsumsq += t[i] * t[i]; }
  const int ls = 1 << level;
  px = px / ls; py = py / ls;
  const unsigned nRange = ((unsigned)range_in + ls - 1) / ls;
  int nTop = py - (int)nRange, nBottomPlusOne = py + (int)nRange + 1, nLeft = px - (int)nRange, nRight = px + (int)nRange;
  *score = max_ssd + 1;
  best_xy[0] = best_xy[1] = 0;
  if (nTop < 0) nTop = 0;
  if (nTop >= h) return 0;
  if (nBottomPlusOne <= 0) return 0;
  if (nLeft < 0) nLeft = 0;
  if (nLeft >= w) return 0;
  int bx = 0, by = 0, nBest = max_ssd + 1;
  if (exhaustive) {
    for (int y = nTop; y < nBottomPlusOne && y < h; y++)
      for (int x = nLeft; x <= nRight && x < w; x++) {
        if ((unsigned)((px - x) * (px - x) + (py - y) * (py - y)) > nRange * nRange) continue;
        const int s = ora_zmssd(im, w, h, stride, t, tsum, tsumsq, x, y, max_ssd);
        if (s < nBest) { bx = x; by = y; nBest = s; }
      }
  } else {
    int i = lut[nTop];
    const int i_end = nBottomPlusOne >= h ? nc : lut[nBottomPlusOne];
    for (; i < i_end; i++) {
      const int x = cxy[2 * i], y = cxy[2 * i + 1];
      if (x < nLeft || x > nRight) continue;
      if ((unsigned)((px - x) * (px - x) + (py - y) * (py - y)) > nRange * nRange) continue;
      const int s = ora_zmssd(im, w, h, stride, t, tsum, tsumsq, x, y, max_ssd);
      if (s < nBest) { bx = x; by = y; nBest = s; }
    }
  }
  *score = nBest;
  if (nBest < max_ssd) { best_xy[0] = bx; best_xy[1] = by; return 1; }
  return 0;
}

/* [3P] TooN::Cholesky<3> (LDL^T, no square roots) and get_inverse() = backsub(Identity) */
static void toon_chol3_inverse(const double* A, double* inv)
{
  double c[9];
  memcpy(c, A, sizeof(c));
  for (int col = 0; col < 3; col++) {
    double inv_diag = 1;
    for (int row = col; row < 3; row++) {
      double val = c[row * 3 + col];
      for (int col2 = 0; col2 < col; col2++) val -= c[col2 * 3 + col] * c[row * 3 + col2];
      if (row == col) { c[row * 3 + col] = val; if (val == 0) goto factored; inv_diag = 1 / val; }
      else { c[col * 3 + row] = val; c[row * 3 + col] = val * inv_diag; }
    }
  }
factored:
  for (int k = 0; k < 3; k++) {
    double v[3] = { 0, 0, 0 }, y[3], r[3];
    v[k] = 1;
    for (int i = 0; i < 3; i++) { double val = v[i]; for (int j = 0; j < i; j++) val -= c[i * 3 + j] * y[j]; y[i] = val; }
    for (int i = 0; i < 3; i++) y[i] /= c[i * 3 + i];
    for (int i = 2; i >= 0; i--) { double val = y[i]; for (int j = i + 1; j < 3; j++) val -= c[j * 3 + i] * r[j]; r[i] = val; }
    inv[0 * 3 + k] = r[0]; inv[1 * 3 + k] = r[1]; inv[2 * 3 + k] = r[2];
  }
}

/* src/PatchFinder.cc:362-470 */
int ora_subpix(const uint8_t* im, int w, int h, int stride, const uint8_t* t, int level, double* pos, int max_its)
{
  float jx[36], jy[36];
  double H[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 }, Hinv[9];
  for (int x = 1; x < 7; x++)                                     /* loop order as in :367-379 */
    for (int y = 1; y < 7; y++) {
      const double gx = 0.5 * (t[y * 8 + x + 1] - t[y * 8 + x - 1]);
      const double gy = 0.5 * (t[(y + 1) * 8 + x] - t[(y - 1) * 8 + x]);
      jx[(y - 1) * 6 + (x - 1)] = (float)gx;
      jy[(y - 1) * 6 + (x - 1)] = (float)gy;
      const double g[3] = { gx, gy, 1.0 };
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) H[r * 3 + c] += g[r] * g[c];
    }
  toon_chol3_inverse(H, Hinv);
  double mean_diff = 0.0;
  const int ls = 1 << level;
  for (int it = 0; it < max_its; it++) {
    const double cxl = (pos[0] + 0.5) / ls - 0.5, cyl = (pos[1] + 0.5) / ls - 0.5;   /* LevelNPos */
    const int rx = (int)(cxl > 0.0 ? cxl + 0.5 : cxl - 0.5), ry = (int)(cyl > 0.0 ? cyl + 0.5 : cyl - 0.5);
    if (!(rx >= 5 && ry >= 5 && rx < w - 5 && ry < h - 5)) return 0;                 /* border 8/2+1 */
    const double bx = cxl - 4, by = cyl - 4;
    double acc[3] = { 0, 0, 0 };
    const double dX = bx - floor(bx), dY = by - floor(by);
    const float fMixTL = (float)((1.0 - dX) * (1.0 - dY));
    const float fMixTR = (float)((dX) * (1.0 - dY));
    const float fMixBL = (float)((1.0 - dX) * (dY));
    const float fMixBR = (float)((dX) * (dY));
    const int ibx = (int)bx, iby = (int)by;
    for (int y = 1; y < 7; y++) {
      const uint8_t* tl = im + (size_t)(iby + y) * stride + ibx + 1;
      for (int x = 1; x < 7; x++) {
        const float fPixel = fMixTL * tl[0] + fMixTR * tl[1] + fMixBL * tl[stride] + fMixBR * tl[stride + 1];
        tl++;
        const double dDiff = fPixel - t[y * 8 + x] + mean_diff;
        acc[0] += dDiff * jx[(y - 1) * 6 + (x - 1)];
        acc[1] += dDiff * jy[(y - 1) * 6 + (x - 1)];
        acc[2] += dDiff;
      }
    }
    double upd[3];
    for (int r = 0; r < 3; r++) upd[r] = Hinv[r * 3 + 0] * acc[0] + Hinv[r * 3 + 1] * acc[1] + Hinv[r * 3 + 2] * acc[2];
    pos[0] -= upd[0] * ls;
    pos[1] -= upd[1] * ls;
    mean_diff -= upd[2];
    const double u2 = upd[0] * upd[0] + upd[1] * upd[1];
    if (u2 < 0) return 0;
    if (u2 < 0.03 * 0.03) return 1;
  }
  return 0;
}

/* src/MiniPatch.cc:34-57 (9x9, mnHalfPatchSize 4, mnMaxSSD 9999) */
int ora_minipatch_ssd(const uint8_t* im, int w, int h, int stride, const uint8_t* patch, int x, int y)
{
  if (!(x >= 4 && y >= 4 && x < w - 4 && y < h - 4)) return 9999 + 1;
  int s = 0;
  for (int r = 0; r < 9; r++) {
    const uint8_t* ip = im + (size_t)(y - 4 + r) * stride + (x - 4);
    for (int c = 0; c < 9; c++) { const int d = ip[c] - patch[r * 9 + c]; s += d * d; }
  }
  return s;
}
/* src/MiniPatch.cc:61-113 */
int ora_minipatch_find(const uint8_t* im, int w, int h, int stride, const uint8_t* patch, const int32_t* cxy, int nc,
                       const int32_t* lut, int n_lut, int range, int32_t* pos)
{
  int bx = 0, by = 0, nBest = 9999 + 1;
  const int tlx = pos[0] - range, tly = pos[1] - range, brx = pos[0] + range, bry = pos[1] + range;
  int i = 0;
  if (!lut) { for (i = 0; i < nc; i++) if (cxy[2 * i + 1] >= tly) break; }
  else {
    int top = tly;
    if (top < 0) top = 0;
    if (top >= n_lut) top = n_lut - 1;
    i = lut[top];
  }
  for (; i < nc; i++) {
    const int x = cxy[2 * i], y = cxy[2 * i + 1];
    if (x < tlx || x > brx) continue;
    if (y > bry) break;
    const int s = ora_minipatch_ssd(im, w, h, stride, patch, x, y);
    if (s < nBest) { bx = x; by = y; nBest = s; }
  }
  if (nBest < 9999) { pos[0] = bx; pos[1] = by; return 1; }
  return 0;
}

/* include/mcptam/TrackerData.h:102-129 (Project, GetDerivsUnsafe) + src/PatchFinder.cc:69-122
   (CalcSearchLevelAndWarpMatrix).  pose_Rt: camera-from-world, 12 doubles.  Returns the search level or -1. */
int ora_project_point(const OraTaylorCam* cam, const double* pose_Rt, const double* pw, const double* right_w, const double* down_w,
                      double* px2, double* derivs4, double* warp_inv4, double* v3cam, int* in_image)
{
  const double* R = pose_Rt; const double* t = pose_Rt + 9;
  double vc[3], mr[3], md[3];
  for (int i = 0; i < 3; i++) {
    vc[i] = R[i * 3] * pw[0] + R[i * 3 + 1] * pw[1] + R[i * 3 + 2] * pw[2] + t[i];
    mr[i] = R[i * 3] * right_w[0] + R[i * 3 + 1] * right_w[1] + R[i * 3 + 2] * right_w[2];
    md[i] = R[i * 3] * down_w[0] + R[i * 3 + 1] * down_w[1] + R[i * 3 + 2] * down_w[2];
  }
  const int invalid = ora_cam_project(cam, vc, px2, derivs4);
  for (int i = 0; i < 3; i++) v3cam[i] = vc[i];
  *in_image = 0;
  if (!invalid && !(px2[0] < 0 || px2[1] < 0 || px2[0] > cam->image_size[0] || px2[1] > cam->image_size[1])) *in_image = 1;   /* note '>' (TrackerData.h:113) */
  double dth[3], dph[3];
  ora_cam_sphere_deriv(vc, dth, dph);
  const double r0 = dth[0] * mr[0] + dth[1] * mr[1] + dth[2] * mr[2], r1 = dph[0] * mr[0] + dph[1] * mr[1] + dph[2] * mr[2];
  const double d0 = dth[0] * md[0] + dth[1] * md[1] + dth[2] * md[2], d1 = dph[0] * md[0] + dph[1] * md[1] + dph[2] * md[2];
  /* mm2WarpInverse.T()[0] = D * right ; .T()[1] = D * down   (row-major out) */
  warp_inv4[0] = derivs4[0] * r0 + derivs4[1] * r1; warp_inv4[2] = derivs4[2] * r0 + derivs4[3] * r1;
  warp_inv4[1] = derivs4[0] * d0 + derivs4[1] * d1; warp_inv4[3] = derivs4[2] * d0 + derivs4[3] * d1;
  double dDet = warp_inv4[0] * warp_inv4[3] - warp_inv4[1] * warp_inv4[2];
  int level = 0;
  while (dDet > 3 && level < 3) { level++; dDet *= 0.25; }
  if (dDet > 3 || dDet < 0.5) return -1;
  return level;
}

/* Batch driver of the per-patch oracle functions (src/Tracker.cc:1299-1377 loop body) for CPU timing without
   per-call Python overhead.  req: n x {src_level, src_cx, src_cy, search_level, pred_x, pred_y, range, subpix_its,
   exhaustive} ints + warp (2x2 m2 already inverted/scaled) doubles.  Returns the number found. */
int ora_search_patches_batch(const uint8_t* const* src_pyr, const uint8_t* const* tgt_pyr, const int* widths, const int* heights,
                             const int32_t* const* corners, const int* n_corners, const int32_t* const* luts, int n,
                             const int32_t* req_i /*9 per req*/, const double* m2 /*4 per req*/, double* found_xy /*2 per req*/,
                             int32_t* found_flag)
{
  int nf = 0;
  for (int i = 0; i < n; i++) {
    const int32_t* r = req_i + 9 * i;
    const int sl = r[0], lvl = r[3];
    uint8_t t[64];
    found_flag[i] = 0;
    if (ora_patch_template(src_pyr[sl], widths[sl], heights[sl], widths[sl], m2 + 4 * i, r[1], r[2], t)) continue;
    int32_t best[2], score;
    if (!ora_find_patch_coarse(tgt_pyr[lvl], widths[lvl], heights[lvl], widths[lvl], corners[lvl], n_corners[lvl], luts[lvl], t, lvl,
                               r[4], r[5], r[6], r[8], best, &score)) continue;
    const int ls = 1 << lvl;
    double pos[2] = { (best[0] + 0.5) * ls - 0.5, (best[1] + 0.5) * ls - 0.5 };
    if (r[7] > 0 && !ora_subpix(tgt_pyr[lvl], widths[lvl], heights[lvl], widths[lvl], t, lvl, pos, r[7])) continue;
    found_xy[2 * i] = pos[0]; found_xy[2 * i + 1] = pos[1];
    found_flag[i] = 1;
    nf++;
  }
  return nf;
}

/* TrackerData::ProjectAndDerivs + CalcJacobian (include/mcptam/TrackerData.h:102-178): 2x6 Jacobian of the pixel
   w.r.t. the MKF base pose.  base_Rt: base-from-world, cfb_Rt: cam-from-base.  J12 row-major 2x6. */
int ora_calc_jacobian(const OraTaylorCam* cam, const double* base_Rt, const double* cfb_Rt, const double* pw, double* px2,
                      double* derivs4, double* J12)
{
  double vb[3], vc[3];
  for (int i = 0; i < 3; i++) vb[i] = base_Rt[i * 3] * pw[0] + base_Rt[i * 3 + 1] * pw[1] + base_Rt[i * 3 + 2] * pw[2] + base_Rt[9 + i];
  for (int i = 0; i < 3; i++) vc[i] = cfb_Rt[i * 3] * vb[0] + cfb_Rt[i * 3 + 1] * vb[1] + cfb_Rt[i * 3 + 2] * vb[2] + cfb_Rt[9 + i];
  const int invalid = ora_cam_project(cam, vc, px2, derivs4);
  double dth[3], dph[3];
  ora_cam_sphere_deriv(vc, dth, dph);
  for (int m = 0; m < 6; m++) {
    double mb[3] = { 0, 0, 0 };
    if (m < 3) mb[m] = 1.0;
    else { const int k = m - 3; mb[(k + 1) % 3] = -vb[(k + 2) % 3]; mb[(k + 2) % 3] = vb[(k + 1) % 3]; }
    double mc[3];
    for (int i = 0; i < 3; i++) mc[i] = cfb_Rt[i * 3] * mb[0] + cfb_Rt[i * 3 + 1] * mb[1] + cfb_Rt[i * 3 + 2] * mb[2];
    const double s0 = dth[0] * mc[0] + dth[1] * mc[1] + dth[2] * mc[2], s1 = dph[0] * mc[0] + dph[1] * mc[1] + dph[2] * mc[2];
    J12[m] = derivs4[0] * s0 + derivs4[1] * s1;
    J12[6 + m] = derivs4[2] * s0 + derivs4[3] * s1;
  }
  return invalid;
}

static int cmp_d(const void* a, const void* b) { const double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
/* Tracker::CalcPoseUpdate (src/Tracker.cc:1386-1511) with [3P] TooN WLS<6> (prior 100, Cholesky LDL^T).
   estimator: 0 Tukey, 1 Cauchy, 2 Huber (include/mcptam/MEstimator.h).  outlier[i] = 1 where the weight is exactly 0. */
int ora_pose_update(int n, const double* found_xy, const double* image_xy, const double* sqrt_inv_noise, const double* jac12,
                    const int32_t* found, int estimator, double override_sigma, double* mu6, double* sigma_sq_out, int32_t* outlier)
{
  double* e2 = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double* ex = (double*)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
  int nv = 0;
  for (int i = 0; i < n; i++) {
    outlier[i] = 0;
    if (!found[i]) continue;
    ex[2 * i] = sqrt_inv_noise[i] * (found_xy[2 * i] - image_xy[2 * i]);
    ex[2 * i + 1] = sqrt_inv_noise[i] * (found_xy[2 * i + 1] - image_xy[2 * i + 1]);
    e2[nv++] = ex[2 * i] * ex[2 * i] + ex[2 * i + 1] * ex[2 * i + 1];
  }
  for (int k = 0; k < 6; k++) mu6[k] = 0;
  *sigma_sq_out = 0;
  if (nv == 0) { free(e2); free(ex); return 0; }
  double sig2;
  if (override_sigma > 0) sig2 = override_sigma;
  else {
    qsort(e2, (size_t)nv, sizeof(double), cmp_d);
    const double med = e2[nv / 2];
    double s = 1.4826 * (1 + 5.0 / (double)((size_t)nv * 2 - 6)) * sqrt(med);
    s = (estimator == 2 ? 1.345 : 4.6851) * s;
    sig2 = s * s;
  }
  *sigma_sq_out = sig2;
  double Cinv[36], vec[6];
  for (int i = 0; i < 36; i++) Cinv[i] = 0;
  for (int i = 0; i < 6; i++) { Cinv[i * 6 + i] = 100.0; vec[i] = 0; }
  int n_in = 0;
  for (int i = 0; i < n; i++) {
    if (!found[i]) continue;
    const double esq = ex[2 * i] * ex[2 * i] + ex[2 * i + 1] * ex[2 * i + 1];
    double w;
    if (estimator == 0) { const double sq = esq > sig2 ? 0.0 : 1.0 - (esq / sig2); w = sq * sq; }
    else if (estimator == 1) w = 1.0 / (1.0 + esq / sig2);
    else w = esq < sig2 ? 1.0 : sqrt(sig2 / esq);
    if (w == 0.0) { outlier[i] = 1; continue; }
    n_in++;
    for (int r = 0; r < 2; r++) {
      double J[6], Jw[6];
      for (int k = 0; k < 6; k++) { J[k] = sqrt_inv_noise[i] * jac12[12 * i + 6 * r + k]; Jw[k] = J[k] * w; }
      for (int a = 0; a < 6; a++) { for (int b = 0; b < 6; b++) Cinv[a * 6 + b] += Jw[a] * J[b]; vec[a] += ex[2 * i + r] * Jw[a]; }
    }
  }
  /* TooN::Cholesky<6> (LDL^T) backsub */
  double c[36];
  for (int i = 0; i < 36; i++) c[i] = Cinv[i];
  for (int col = 0; col < 6; col++) {
    double inv_diag = 1;
    for (int row = col; row < 6; row++) {
      double val = c[row * 6 + col];
      for (int col2 = 0; col2 < col; col2++) val -= c[col2 * 6 + col] * c[row * 6 + col2];
      if (row == col) { c[row * 6 + col] = val; inv_diag = 1 / val; }
      else { c[col * 6 + row] = val; c[row * 6 + col] = val * inv_diag; }
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) { double val = vec[i]; for (int j = 0; j < i; j++) val -= c[i * 6 + j] * y[j]; y[i] = val; }
  for (int i = 0; i < 6; i++) y[i] /= c[i * 6 + i];
  for (int i = 5; i >= 0; i--) { double val = y[i]; for (int j = i + 1; j < 6; j++) val -= c[j * 6 + i] * mu6[j]; mu6[i] = val; }
  free(e2); free(ex);
  return n_in;
}

/* ------------------------------------------------------------------------------------------
 * KeyFrame::MakeKeyFrame_Rest candidate generation (src/KeyFrame.cc:363-531)
 * ------------------------------------------------------------------------------------------ */
/* [3P] libCVD fast_corner.cpp old_style_corner_score (the score fast_nonmax suppresses on):
   sp = sum over the ring of (p - (c + b)) where p > c + b,  sn = sum of ((c - b) - p) where p < c - b; max(sp, sn). */
int ora_fast_old_score(const uint8_t* im, int stride, int x, int y, int barrier)
{
  const uint8_t* p = im + (size_t)y * stride + x;
  const int cb = *p + barrier, c_b = *p - barrier;
  int sp = 0, sn = 0;
  for (int i = 0; i < 16; i++) {
    const int v = p[RING_DY[i] * stride + RING_DX[i]];
    if (v > cb) sp += v - cb;
    else if (v < c_b) sn += c_b - v;
  }
  return sp > sn ? sp : sn;
}

/* [3P] libCVD fast_nonmax (src/KeyFrame.cc:393,411): 3x3 suppression among the listed corners on the old-style score.
   strict = 0 (libCVD nonmax_suppression): a corner is dropped when a neighbouring corner scores strictly higher;
   strict = 1 (nonmax_suppression_strict): dropped when a neighbour scores higher or equal.
   Writes 1/0 per input corner into keep[]; returns the number kept. */
int ora_fast_nonmax(const uint8_t* im, int w, int h, int stride, const int32_t* cxy, int nc, int barrier, int strict, uint8_t* keep)
{
  int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++) map[i] = -1;
  for (int i = 0; i < nc; i++) map[(size_t)cxy[2 * i + 1] * w + cxy[2 * i]] = ora_fast_old_score(im, stride, cxy[2 * i], cxy[2 * i + 1], barrier);
  int n = 0;
  for (int i = 0; i < nc; i++) {
    const int x = cxy[2 * i], y = cxy[2 * i + 1], s = map[(size_t)y * w + x];
    int ok = 1;
    for (int dy = -1; dy <= 1 && ok; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        if (!dx && !dy) continue;
        const int xx = x + dx, yy = y + dy;
        if (xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
        const int o = map[(size_t)yy * w + xx];
        if (o < 0) continue;
        if (strict ? (o >= s) : (o > s)) { ok = 0; break; }
      }
    keep[i] = (uint8_t)ok;
    n += ok;
  }
  free(map);
  return n;
}

typedef struct { double score; int32_t x, y; } ora_cand_t;
/* std::sort(rbegin, rend) on std::pair<double, CVD::ImageRef>: descending score, ties by descending ImageRef
   (CVD::ImageRef::operator< is raster order: y, then x). */
static int cmp_cand_desc(const void* a, const void* b)
{
  const ora_cand_t* p = (const ora_cand_t*)a; const ora_cand_t* q = (const ora_cand_t*)b;
  if (p->score != q->score) return p->score > q->score ? -1 : 1;
  if (p->y != q->y) return p->y > q->y ? -1 : 1;
  if (p->x != q->x) return p->x > q->x ? -1 : 1;
  return 0;
}

/* One level of MakeKeyFrame_Rest.  cxy/nc/lut: Level::vCorners + vCornerRowLUT of the current image; fast_thresh: nFastThresh.
   prev_*: imagePrev[0] / vCornersPrev[0] (NULL when the history is empty), n_prev = imagePrev.size().
   Returns the number of candidates, written to out_xy / out_score; *n_max = vScoresAndMaxCorners.size(). */
int ora_keyframe_rest_level(const uint8_t* im, int w, int h, int stride, const int32_t* cxy, int nc, const int32_t* lut, int fast_thresh,
                            int use_shi, int use_thresh, double top_fraction, double thresh, int nonmax_strict,
                            const uint8_t* prev_im, const int32_t* prev_cxy, int prev_nc, const int32_t* prev_lut, int n_prev,
                            int32_t* out_xy, double* out_score, int cap, int32_t* n_max)
{
  uint8_t* keep = (uint8_t*)malloc((size_t)nc + 1);
  ora_fast_nonmax(im, w, h, stride, cxy, nc, fast_thresh, nonmax_strict, keep);
  ora_cand_t* v = (ora_cand_t*)malloc(sizeof(ora_cand_t) * ((size_t)nc + 1));
  int nv = 0;
  for (int i = 0; i < nc; i++) {
    if (!keep[i]) continue;
    const int x = cxy[2 * i], y = cxy[2 * i + 1];
    if (!(x >= 10 && y >= 10 && x < w - 10 && y < h - 10)) continue;          /* in_image_with_border(.., 10), :400,415 */
    double sc;
    if (use_shi) sc = ora_shitomasi(im, stride, 3, x, y);
    else { int32_t s; const int32_t one[2] = { x, y }; ora_fast10_score_bisect(im, stride, one, 1, fast_thresh, &s); sc = s; }
    v[nv].score = sc; v[nv].x = x; v[nv].y = y; nv++;
  }
  if (n_max) *n_max = nv;
  ora_cand_t* cand = (ora_cand_t*)malloc(sizeof(ora_cand_t) * ((size_t)nv + 1));
  int ncand = 0;
  if (!use_thresh) {
    qsort(v, (size_t)nv, sizeof(ora_cand_t), cmp_cand_desc);
    const int n_use = (int)(nv * top_fraction);
    for (int i = 0; i < n_use && i < nv; i++) cand[ncand++] = v[i];
  } else {
    for (int i = 0; i < nv; i++) if (v[i].score > thresh) cand[ncand++] = v[i];
  }
  int nout = 0;
  for (int i = 0; i < ncand; i++) {
    if (prev_im && n_prev > 0) {                                               /* :455-527 stable-point pruning */
      uint8_t patch[81];
      const int cx = cand[i].x, cy = cand[i].y;
      for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) patch[r * 9 + c] = im[(size_t)(cy - 4 + r) * stride + cx - 4 + c];
      int32_t pp[2] = { cx, cy };
      if (!ora_minipatch_find(prev_im, w, h, stride, patch, prev_cxy, prev_nc, prev_lut, h, n_prev * 10, pp)) continue;
      for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) patch[r * 9 + c] = prev_im[(size_t)(pp[1] - 4 + r) * stride + pp[0] - 4 + c];
      int32_t pn[2] = { pp[0], pp[1] };
      if (!ora_minipatch_find(im, w, h, stride, patch, cxy, nc, lut, h, n_prev * 10, pn)) continue;
      const int dx = pn[0] - cx, dy = pn[1] - cy;
      if (dx * dx + dy * dy > 2) continue;
    }
    if (nout < cap) { out_xy[2 * nout] = cand[i].x; out_xy[2 * nout + 1] = cand[i].y; out_score[nout] = cand[i].score; }
    nout++;
  }
  free(keep); free(v); free(cand);
  return nout;
}
