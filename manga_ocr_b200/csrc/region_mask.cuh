// Crop staging on the device: the polygon mask of a selection,
// rasterised exactly like the reference's
//     mask = np.zeros(crop.shape[:2], np.uint8); cv2.fillPoly(mask, [polygon - bbox.topLeft], 255)
// (reference/src/ui/main_window.py:6499-6502; OpenCV 4.x drawing.cpp: CollectPolyEdges +
// FillEdgeCollection + Line/LineIterator + clipLine, lineType = LINE_8, shift = 0).
//
// Split of the work: the host turns the n polygon points into n fill edges and n clipped line
// segments (O(points), including clipLine's double arithmetic, so the device never touches floating
// point); the device rasterises - one warp per scanline for the even-odd fill, one thread per
// segment for the outline - and the preprocess kernel applies the mask while it reads the page
// (`inside ? pixel : 255`, the reference's bitwise_and / add composite on white, :6503-6506).
#pragma once
#include <stdint.h>

#include <vector>

#include "common.cuh"

namespace mocr {

constexpr int kMaskMaxEdges = 1024;      // polygon points per selection
constexpr int kMaskRowsPerCta = 8;       // one warp per scanline
constexpr int kXyShift = 16;
constexpr long long kXyHalf = 1ll << 15;

struct MaskEdge {          // PolyEdge: active on scanlines [y0, y1), x advances by dx per scanline
  int y0, y1;
  long long x, dx;         // 48.16 fixed point; x carries the +0.5 of the non-antialiased path
};
struct MaskLine {          // outline segment, already clipped to the mask rectangle
  int x1, y1, x2, y2;
};
struct MaskJob {
  long long mask_off;      // byte offset of the [h][w] uint8 mask in the mask arena
  int h, w;
  int edge0, n_edges;      // fill edges (0: nothing to fill)
  int line0, n_lines;
};

// cv::clipLine(Size(w, h), pt1, pt2), including the partial updates it leaves when the segment is outside.
inline bool clip_line_cv(int w, int h, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += static_cast<long long>(static_cast<double>(a - y1) * (x2 - x1) / (y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += static_cast<long long>(static_cast<double>(a - y2) * (x2 - x1) / (y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += static_cast<long long>(static_cast<double>(a - x1) * (y2 - y1) / (x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += static_cast<long long>(static_cast<double>(a - x2) * (y2 - y1) / (x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// CollectPolyEdges for one polygon (points relative to the mask origin): appends the outline
// segments that survive clipping and the fill edges; returns false when nothing is to be filled
// (fewer than two edges, or FillEdgeCollection's bounding-box early exit).
inline bool collect_poly_edges(int h, int w, const int32_t* pts, int n, std::vector<MaskEdge>& edges, std::vector<MaskLine>& lines) {
  const size_t e_begin = edges.size();
  if (n <= 0) return false;
  long long p0x = pts[2 * (n - 1)], p0y = pts[2 * (n - 1) + 1];
  for (int i = 0; i < n; ++i) {
    const long long p1x = pts[2 * i], p1y = pts[2 * i + 1];
    long long t0x = p0x, t0y = p0y, t1x = p1x, t1y = p1y;
    const bool inside = p0x >= 0 && p0x < w && p1x >= 0 && p1x < w && p0y >= 0 && p0y < h && p1y >= 0 && p1y < h;
    bool visible = true;
    if (!inside) visible = clip_line_cv(w, h, t0x, t0y, t1x, t1y);
    if (visible) lines.push_back(MaskLine{static_cast<int>(t0x), static_cast<int>(t0y), static_cast<int>(t1x), static_cast<int>(t1y)});
    // slope and x from the clipped segment; y only when that segment is not horizontal
    long long c0y = p0y, c1y = p1y;
    if (!inside && t0y != t1y) { c0y = t0y; c1y = t1y; }
    if (p0y != p1y) {
      const long long f0 = t0x * (1ll << kXyShift) + kXyHalf, f1 = t1x * (1ll << kXyShift) + kXyHalf;
      MaskEdge e;
      e.dx = (f1 - f0) / (c1y - c0y);                  // C++ division truncates toward zero, as in OpenCV
      if (p0y < p1y) { e.y0 = static_cast<int>(p0y); e.y1 = static_cast<int>(p1y); e.x = f0 + (p0y - c0y) * e.dx; }
      else { e.y0 = static_cast<int>(p1y); e.y1 = static_cast<int>(p0y); e.x = f1 + (p1y - c1y) * e.dx; }
      edges.push_back(e);
    }
    p0x = p1x;
    p0y = p1y;
  }
  const size_t ne = edges.size() - e_begin;
  bool fill = ne >= 2;
  if (fill) {
    long long y_min = INT32_MAX, y_max = INT32_MIN, x_min = INT64_MAX, x_max = INT64_MIN;
    for (size_t k = e_begin; k < edges.size(); ++k) {
      const MaskEdge& e = edges[k];
      const long long xe = e.x + static_cast<long long>(e.y1 - e.y0) * e.dx;
      y_min = std::min<long long>(y_min, e.y0);
      y_max = std::max<long long>(y_max, e.y1);
      x_min = std::min(x_min, std::min(e.x, xe));
      x_max = std::max(x_max, std::max(e.x, xe));
    }
    if (y_max < 0 || y_min >= h || x_max < 0 || x_min >= (static_cast<long long>(w) << kXyShift)) fill = false;
  }
  if (!fill) edges.resize(e_begin);
  return fill;
}

// Even-odd fill.  A pixel p of scanline y is inside iff some consecutive pair (a, b) of the sorted
// crossings has a <= p <= b (FillEdgeCollection draws [ceil(a), floor(b)]).  With S = #crossings < p
// and E = #crossings == p that is (S odd) || (E > 0), which needs no sort: each crossing t (in 1/65536
// pixel, relative to the pixel centre grid) becomes the key 2*floor(t) + (t is fractional), so
// t < p  <=>  key < 2p  and  t == p  <=>  key == 2p.
// grid = (ceil(max_h / kMaskRowsPerCta), jobs), block = 32 * kMaskRowsPerCta.
__global__ void __launch_bounds__(32 * kMaskRowsPerCta)
region_mask_fill_kernel(const MaskJob* __restrict__ jobs, const MaskEdge* __restrict__ edges, uint8_t* __restrict__ arena) {
  __shared__ int keys[kMaskRowsPerCta][kMaskMaxEdges];
  const MaskJob job = jobs[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = blockIdx.x * kMaskRowsPerCta + warp;
  if (y >= job.h) return;
  int* k = keys[warp];
  int n_active = 0;
  for (int e0 = 0; e0 < job.n_edges; e0 += 32) {
    const int e = e0 + lane;
    bool act = false;
    int key = 0;
    if (e < job.n_edges) {
      const MaskEdge ed = edges[job.edge0 + e];
      if (ed.y0 <= y && y < ed.y1) {
        act = true;
        long long t = ed.x + static_cast<long long>(y - ed.y0) * ed.dx - kXyHalf;
        const long long lim = 1ll << 44;                 // far outside any mask: order against the pixel range is kept
        t = t < -lim ? -lim : (t > lim ? lim : t);
        const long long kk = (t >> kXyShift) * 2 + ((t & 0xffff) != 0 ? 1 : 0);
        key = static_cast<int>(kk < -(1ll << 30) ? -(1ll << 30) : (kk > (1ll << 30) ? (1ll << 30) : kk));
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (act) k[n_active + __popc(m & ((1u << lane) - 1u))] = key;
    n_active += __popc(m);
  }
  __syncwarp();
  uint8_t* row = arena + job.mask_off + static_cast<long long>(y) * job.w;
  for (int p = lane; p < job.w; p += 32) {
    int s = 0, eq = 0;
    for (int i = 0; i < n_active; ++i) {
      const int key = k[i];
      s += key < 2 * p;
      eq |= key == 2 * p;
    }
    row[p] = ((s & 1) | eq) ? 255 : 0;
  }
}

// Outline: cv::line (LINE_8) = LineIterator walked left to right over the clipped segment.
// One thread per segment; grid = ceil(total_lines / 128).
__global__ void __launch_bounds__(128)
region_mask_lines_kernel(const MaskJob* __restrict__ jobs, int n_jobs, const MaskLine* __restrict__ lines, int n_lines, uint8_t* __restrict__ arena) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= n_lines) return;
  int j = 0;                                   // the job that owns this segment (jobs hold consecutive line ranges)
  while (j + 1 < n_jobs && li >= jobs[j + 1].line0) ++j;
  const MaskJob job = jobs[j];
  const MaskLine ln = lines[li];
  int x1 = ln.x1, y1 = ln.y1, dx = ln.x2 - ln.x1, dy = ln.y2 - ln.y1, sy = 1;
  if (dx < 0) { dx = -dx; dy = -dy; x1 = ln.x2; y1 = ln.y2; }
  if (dy < 0) { dy = -dy; sy = -1; }
  const bool vert = dy > dx;
  if (vert) { const int t = dx; dx = dy; dy = t; }
  int err = dx - (dy + dy);
  const int plus = dx + dx, minus = -(dy + dy);
  uint8_t* mask = arena + job.mask_off;
  int x = x1, y = y1;
  for (int i = 0; i <= dx; ++i) {
    mask[static_cast<long long>(y) * job.w + x] = 255;
    const bool neg = err < 0;
    err += minus + (neg ? plus : 0);
    if (vert) { y += sy; x += neg ? 1 : 0; }
    else { x += 1; y += neg ? sy : 0; }
  }
}

}  // namespace mocr
