// Host-side finishing of one component: cv2.boxPoints(cv2.minAreaRect(contour)) (utils.py:56-57).
//
// OpenCV is a third-party dependency of the reference (opencv-python>=3.4,<4.0, requirements.txt:5),
// not part of its tree; this restates its published algorithm: convex hull of the point set, then
// float32 rotating calipers over the hull edges keeping the LAST rectangle of minimal area
// (area <= minarea), then the RotatedRect -> 4 corner conversion.  The hull is fed in the order
// OpenCV's convexHull produces for a findContours contour: counter-clockwise in (x right, y up)
// terms, ending at the contour's first point (top-most, then left-most pixel).  Against
// cv2 4.13 on 5,795 contours the rounded x4 boxes agree as corner sets in 99.9 % of cases; the
// rest are exact equal-area ties (documented in DESIGN.md, compared as ties in the tests).
//
// Attribution: rotating_calipers() below follows the structure, float32 operation order and tie rule of
// rotatingCalipers() in OpenCV's modules/imgproc/src/rotcalipers.cpp (Copyright (C) 2000, Intel Corporation;
// Copyright (C) OpenCV contributors; OpenCV 4.5+ is distributed under the Apache License, Version 2.0,
// http://www.apache.org/licenses/LICENSE-2.0 - earlier releases under the 3-clause BSD license), because the boxes must
// be bit-identical to cv2.minAreaRect's.  ccl_boxes_kernel in ubd_ccl.cuh is the same procedure, one warp per component.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/ubd.h"

namespace {

struct Pt { int x, y; };

inline long long cross(const Pt& o, const Pt& a, const Pt& b) {
  return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// Strict convex hull (no collinear points), Andrew monotone chain, exact integer arithmetic.
void convex_hull(std::vector<Pt>& pts, std::vector<Pt>& hull) {
  std::sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  pts.erase(std::unique(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) { return a.x == b.x && a.y == b.y; }),
            pts.end());
  hull.clear();
  const int n = (int)pts.size();
  if (n <= 2) { hull = pts; return; }
  static thread_local std::vector<Pt> h;                 // scratch reused across calls (one call per component)
  h.resize(2 * n);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; --i) {
    while (k >= t && cross(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  h.resize(k - 1);
  if ((int)h.size() <= 2) { hull = h; return; }
  // rotate so that the hull ENDS at the top-most, then left-most vertex (the contour's first point)
  int s = 0;
  for (int i = 1; i < (int)h.size(); ++i)
    if (h[i].y < h[s].y || (h[i].y == h[s].y && h[i].x < h[s].x)) s = i;
  hull.resize(h.size());
  const int m = (int)h.size();
  for (int i = 0; i < m; ++i) hull[i] = h[(s + 1 + i) % m];
}

struct P2f { float x, y; };

// Rotating calipers, minimal-area rectangle: out[0] = corner, out[1], out[2] = edge vectors.
void rotating_calipers(const P2f* points, int n, P2f out[3]) {
  float minarea = 3.402823466e+38f;
  static thread_local std::vector<float> inv_vect_length;
  static thread_local std::vector<P2f> vect;
  inv_vect_length.resize(n);
  vect.resize(n);
  int left = 0, bottom = 0, right = 0, top = 0;
  int seq[4];
  float orientation = 0.f, base_a, base_b = 0.f;
  P2f pt0 = points[0];
  float left_x = pt0.x, right_x = pt0.x, top_y = pt0.y, bottom_y = pt0.y;
  for (int i = 0; i < n; ++i) {
    if (pt0.x < left_x) { left_x = pt0.x; left = i; }
    if (pt0.x > right_x) { right_x = pt0.x; right = i; }
    if (pt0.y > top_y) { top_y = pt0.y; top = i; }
    if (pt0.y < bottom_y) { bottom_y = pt0.y; bottom = i; }
    const P2f pt = points[(i + 1 < n) ? i + 1 : 0];
    const double dx = (double)pt.x - (double)pt0.x, dy = (double)pt.y - (double)pt0.y;
    vect[i].x = (float)dx; vect[i].y = (float)dy;
    inv_vect_length[i] = (float)(1. / sqrt(dx * dx + dy * dy));
    pt0 = pt;
  }
  {
    double ax = vect[n - 1].x, ay = vect[n - 1].y;
    for (int i = 0; i < n; ++i) {
      const double bx = vect[i].x, by = vect[i].y;
      const double convexity = ax * by - ay * bx;
      if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
      ax = bx; ay = by;
    }
  }
  base_a = orientation;
  seq[0] = bottom; seq[1] = right; seq[2] = top; seq[3] = left;
  int best_left = 0, best_bottom = 0;
  float best_a = 1.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
  for (int k = 0; k < n; ++k) {
    const float dp[4] = {
        +base_a * vect[seq[0]].x + base_b * vect[seq[0]].y,
        -base_b * vect[seq[1]].x + base_a * vect[seq[1]].y,
        -base_a * vect[seq[2]].x - base_b * vect[seq[2]].y,
        +base_b * vect[seq[3]].x - base_a * vect[seq[3]].y,
    };
    float maxcos = dp[0] * inv_vect_length[seq[0]];
    int main_element = 0;
    for (int i = 1; i < 4; ++i) {
      const float cosalpha = dp[i] * inv_vect_length[seq[i]];
      if (cosalpha > maxcos) { main_element = i; maxcos = cosalpha; }
    }
    {
      const int pindex = seq[main_element];
      const float lead_x = vect[pindex].x * inv_vect_length[pindex];
      const float lead_y = vect[pindex].y * inv_vect_length[pindex];
      switch (main_element) {
        case 0: base_a = lead_x; base_b = lead_y; break;
        case 1: base_a = lead_y; base_b = -lead_x; break;
        case 2: base_a = -lead_x; base_b = -lead_y; break;
        default: base_a = -lead_y; base_b = lead_x; break;
      }
    }
    seq[main_element] += 1;
    if (seq[main_element] == n) seq[main_element] = 0;
    float dx = points[seq[1]].x - points[seq[3]].x;
    float dy = points[seq[1]].y - points[seq[3]].y;
    const float width = dx * base_a + dy * base_b;
    dx = points[seq[2]].x - points[seq[0]].x;
    dy = points[seq[2]].y - points[seq[0]].y;
    const float height = -dx * base_b + dy * base_a;
    const float area = width * height;
    if (area <= minarea) {
      minarea = area;
      best_left = seq[3]; best_a = base_a; best_w = width; best_b = base_b; best_h = height;
      best_bottom = seq[0];
    }
  }
  const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
  const float C1 = A1 * points[best_left].x + points[best_left].y * B1;
  const float C2 = A2 * points[best_bottom].x + points[best_bottom].y * B2;
  const float idet = 1.f / (A1 * B2 - A2 * B1);
  out[0].x = (C1 * B2 - C2 * B1) * idet;
  out[0].y = (A1 * C2 - A2 * C1) * idet;
  out[1].x = A1 * best_w; out[1].y = B1 * best_w;
  out[2].x = A2 * best_h; out[2].y = B2 * best_h;
}

}  // namespace

// RotatedRect (angle in radians as minAreaRect leaves it before the degree conversion) -> cv2.boxPoints corners.
void ubd_rect_to_box(float cx, float cy, float w, float hgt, float angle, float* box) {
  angle = (float)(angle * 180 / 3.1415926535897932384626433832795);
  // RotatedRect::points
  const double a_ = angle * 3.1415926535897932384626433832795 / 180.;
  const float b = (float)cos(a_) * 0.5f;
  const float a = (float)sin(a_) * 0.5f;
  box[0] = cx - a * hgt - b * w;
  box[1] = cy + b * hgt - a * w;
  box[2] = cx + a * hgt - b * w;
  box[3] = cy - b * hgt - a * w;
  box[4] = 2 * cx - box[0];
  box[5] = 2 * cy - box[1];
  box[6] = 2 * cx - box[2];
  box[7] = 2 * cy - box[3];
}

// The rectangle of a component whose hull and calipers were computed on the GPU (ccl_boxes_kernel): ax, ay = the
// first edge vector (out[1] of rotatingCalipers); n_hull <= 2: the degenerate cases from the hull's end points.
void ubd_box_from_device(float cx, float cy, float w, float hgt, float ax, float ay, int n_hull,
                         int x0, int y0, int x1, int y1, float* box) {
  float angle = 0.f;
  if (n_hull > 2) {
    angle = (float)atan2((double)ay, (double)ax);
  } else if (n_hull == 2) {
    if (x1 < x0 || (x1 == x0 && y1 < y0)) { std::swap(x0, x1); std::swap(y0, y1); }     // the host hull lists them by (x, y)
    cx = ((float)x0 + (float)x1) * 0.5f;
    cy = ((float)y0 + (float)y1) * 0.5f;
    const double dx = (double)x1 - x0, dy = (double)y1 - y0;
    w = (float)sqrt(dx * dx + dy * dy);
    hgt = 0.f;
    angle = (float)atan2(dy, dx);
  } else {
    cx = (float)x0; cy = (float)y0; w = hgt = 0.f;
  }
  ubd_rect_to_box(cx, cy, w, hgt, angle, box);
}

// pts: any superset of the component's hull vertices (e.g. its row-run end points).
extern "C" int ubd_min_area_box(const int32_t* pts_xy, int n_pts, float* box) {
  if (box == nullptr || (n_pts > 0 && pts_xy == nullptr) || n_pts < 0) return UBD_ERR_ARG;
  static thread_local std::vector<Pt> pts, hull;
  pts.resize(n_pts);
  for (int i = 0; i < n_pts; ++i) { pts[i].x = pts_xy[2 * i]; pts[i].y = pts_xy[2 * i + 1]; }
  convex_hull(pts, hull);
  const int n = (int)hull.size();
  float cx = 0.f, cy = 0.f, w = 0.f, hgt = 0.f, angle = 0.f;
  if (n > 2) {
    static thread_local std::vector<P2f> hp;
    hp.resize(n);
    for (int i = 0; i < n; ++i) { hp[i].x = (float)hull[i].x; hp[i].y = (float)hull[i].y; }
    P2f out[3];
    rotating_calipers(hp.data(), n, out);
    cx = out[0].x + (out[1].x + out[2].x) * 0.5f;
    cy = out[0].y + (out[1].y + out[2].y) * 0.5f;
    w = (float)sqrt((double)out[1].x * out[1].x + (double)out[1].y * out[1].y);
    hgt = (float)sqrt((double)out[2].x * out[2].x + (double)out[2].y * out[2].y);
    angle = (float)atan2((double)out[1].y, (double)out[1].x);
  } else if (n == 2) {
    cx = ((float)hull[0].x + (float)hull[1].x) * 0.5f;
    cy = ((float)hull[0].y + (float)hull[1].y) * 0.5f;
    const double dx = (double)hull[1].x - hull[0].x, dy = (double)hull[1].y - hull[0].y;
    w = (float)sqrt(dx * dx + dy * dy);
    angle = (float)atan2(dy, dx);
  } else if (n == 1) {
    cx = (float)hull[0].x; cy = (float)hull[0].y;
  }
  ubd_rect_to_box(cx, cy, w, hgt, angle, box);
  return UBD_OK;
}
