// Target rasterisation of the BEV dataset generator:
//   draw_boxes   generating-dataset/generating_train_bev.py:127-139
//     corners = box.bottom_corners(); corners_voxel = car_to_voxel_coords(corners, im.shape, voxel_size, z_offset)
//     cv2.drawContours(im, np.int0([corners_voxel[:, :2]]), 0, (c, c, c), -1)      c = classes.index(box.name) + 1
// Boxes are painted in list order, later boxes over earlier ones, as filled polygons of four integer
// vertices.  The fill rule is OpenCV's (drawing.cpp: CollectPolyEdges + FillEdgeCollection, Line, clipLine,
// LineIterator) as restated and pinned against cv2 4.13 in oracle/draw_oracle.py; this file evaluates the same
// rule per pixel in closed form instead of walking lines and scanlines:
//   * a pixel lies on the Bresenham walk of a clipped edge iff, i major steps from the start, its minor offset is
//     k(i) = floor((2*d*i + D - 1) / (2*D))      (err0 = D - 2d, a minor step whenever err < 0);
//   * a pixel of row y is filled iff it lies in [ceil(xa), floor(xb)] of a pair (0,1) / (2,3) of the row's active
//     scanline edges sorted by their 16.16 fixed-point x = x0 + (y - y0) * dx.
//   DB1 db_setup   one thread per box: fp64 voxel-space affine of the four corners (un-fused multiply, add - the
//                  BEV rule), C truncation, clipLine per edge, line and scanline-edge records
//   DB2 db_raster  one CTA per 8 x 32 pixel tile: the boxes whose vertex bounding box meets the tile are collected in
//                  shared memory (descending order), then one thread per pixel walks them - the first that covers
//                  the pixel wins; every byte of the (H, W) uint8 target is written exactly once (0 = background)
#include "lv_common.cuh"

namespace {

constexpr int DB_THREADS = 256;
constexpr int DB_XY_SHIFT = 16;
constexpr long long DB_COORD_LIMIT = 1ll << 20;   // vertices are clamped to +-2^20 pixels (int64 headroom of x0 + rows*dx)

struct DbLine { int valid, xs, ys, D, d, sy, vert, pad; };
struct DbEdge { int y0, y1; long long x, dx; };
struct DbBox {
  int color, vy_min, vy_max, fy0, fy1, vx_min, vx_max, pad;
  DbLine line[4];
  DbEdge edge[4];
};

struct DbParams {
  const double* corners;     // (n_boxes, 3, 4) car-space bottom corners
  const int32_t* colors;     // (n_boxes)
  const int64_t* box_off;    // (n_frames + 1)
  DbBox* recs;
  uint8_t* target;           // (n_frames, H, W)
  int H, W;
  double m0, m1, t0, t1;
  int64_t n_boxes;
};

// cv::clipLine(Size2l, Point2l&, Point2l&): the points may change even when the result is false
__device__ bool db_clip_line(long long w, long long h, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = w - 1, bottom = h - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

__device__ __forceinline__ long long db_trunc_clamp(double u) {
  // np.int0 / np.intp: C truncation; non-finite and far-away values are clamped (documented limit)
  if (!(u > -(double)DB_COORD_LIMIT)) return -DB_COORD_LIMIT;
  if (!(u < (double)DB_COORD_LIMIT)) return DB_COORD_LIMIT;
  return (long long)u;
}

__global__ void __launch_bounds__(DB_THREADS) db_setup_kernel(DbParams p) {
  const int64_t b = (int64_t)blockIdx.x * DB_THREADS + threadIdx.x;
  if (b >= p.n_boxes) return;
  const double* c = p.corners + b * 12;
  long long vx[4], vy[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // car_to_voxel_coords (generating_train_bev.py:73-82): tm . [x; y; z; 1], tm diagonal + translation
    vx[k] = db_trunc_clamp(__dadd_rn(__dmul_rn(p.m0, c[k]), p.t0));
    vy[k] = db_trunc_clamp(__dadd_rn(__dmul_rn(p.m1, c[4 + k]), p.t1));
  }
  DbBox r;
  r.color = p.colors[b];
  r.pad = 0;
  long long ymn = vy[0], ymx = vy[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    ymn = vy[k] < ymn ? vy[k] : ymn;
    ymx = vy[k] > ymx ? vy[k] : ymx;
  }
  r.vy_min = (int)(ymn < 0 ? 0 : ymn);
  r.vy_max = (int)(ymx > p.H - 1 ? p.H - 1 : ymx);
  long long xmn = vx[0], xmx = vx[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) {
    xmn = vx[k] < xmn ? vx[k] : xmn;
    xmx = vx[k] > xmx ? vx[k] : xmx;
  }
  // every painted pixel lies inside the vertex bounding box (clipped end points stay on their segment, the
  // truncated slopes never overshoot a vertex on an image row); one pixel of slack on either side
  r.vx_min = (int)(xmn - 1 < 0 ? 0 : xmn - 1);
  r.vx_max = (int)(xmx + 1 > p.W - 1 ? p.W - 1 : xmx + 1);
  long long fy0 = (1ll << 40), fy1 = -(1ll << 40);
  int n_edges = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p0x = vx[(i + 3) & 3], p0y = vy[(i + 3) & 3], p1x = vx[i], p1y = vy[i];
    const bool outside = !(p0x >= 0 && p0x < p.W && p1x >= 0 && p1x < p.W && p0y >= 0 && p0y < p.H && p1y >= 0 && p1y < p.H);
    long long q0x = p0x, q0y = p0y, q1x = p1x, q1y = p1y;
    bool visible = true;
    if (outside) visible = db_clip_line(p.W, p.H, q0x, q0y, q1x, q1y);
    // ---- Line(): LineIterator over the clipped segment, left to right
    DbLine& L = r.line[i];
    L.valid = visible ? 1 : 0;
    L.pad = 0;
    {
      long long x1 = q0x, y1 = q0y, dx = q1x - q0x, dy = q1y - q0y;
      int sy = 1;
      if (dx < 0) { dx = -dx; dy = -dy; x1 = q1x; y1 = q1y; }
      if (dy < 0) { dy = -dy; sy = -1; }
      const bool vert = dy > dx;
      L.xs = (int)x1; L.ys = (int)y1;
      L.D = (int)(vert ? dy : dx);
      L.d = (int)(vert ? dx : dy);
      L.sy = sy;
      L.vert = vert ? 1 : 0;
    }
    // ---- scanline edge (CollectPolyEdges): clipped x's (and y's unless the clipped segment is horizontal)
    // define dx; the start is extrapolated back to the unclipped first row
    DbEdge& E = r.edge[i];
    E.y0 = E.y1 = 0; E.x = 0; E.dx = 0;
    if (p0y != p1y) {
      long long c0x = p0x, c0y = p0y, c1x = p1x, c1y = p1y;
      if (outside) {
        c0x = q0x; c1x = q1x;
        if (q0y != q1y) { c0y = q0y; c1y = q1y; }
      }
      const long long dxf = ((c1x - c0x) << DB_XY_SHIFT) / (c1y - c0y);   // C division truncates toward zero
      E.dx = dxf;
      if (p0y < p1y) {
        E.y0 = (int)p0y; E.y1 = (int)p1y;
        E.x = (c0x << DB_XY_SHIFT) + (p0y - c0y) * dxf;
      } else {
        E.y0 = (int)p1y; E.y1 = (int)p0y;
        E.x = (c1x << DB_XY_SHIFT) + (p1y - c1y) * dxf;
      }
      fy0 = E.y0 < fy0 ? E.y0 : fy0;
      fy1 = E.y1 > fy1 ? E.y1 : fy1;
      ++n_edges;
    }
  }
  if (n_edges < 2) { fy0 = 0; fy1 = 0; }
  r.fy0 = (int)(fy0 < 0 ? 0 : fy0);
  r.fy1 = (int)(fy1 > p.H ? p.H : fy1);
  if (r.fy1 < r.fy0) r.fy1 = r.fy0;
  p.recs[b] = r;
}

__device__ __forceinline__ bool db_on_line(const DbLine& L, int x, int y) {
  if (!L.valid) return false;
  const int i = L.vert ? (y - L.ys) * L.sy : x - L.xs;
  if (i < 0 || i > L.D) return false;
  const int k = L.D > 0 ? (int)((2ll * L.d * i + L.D - 1) / (2ll * L.D)) : 0;
  return L.vert ? (x == L.xs + k) : (y == L.ys + L.sy * k);
}

__device__ __forceinline__ bool db_filled(const DbBox& r, int x, int y) {
  if (y < r.fy0 || y >= r.fy1) return false;
  long long xs[4];
  int n = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const DbEdge& E = r.edge[e];
    if (E.y0 <= y && y < E.y1) {
      const long long v = E.x + (long long)(y - E.y0) * E.dx;
      // insertion into the sorted list (at most four entries)
      int j = n++;
#pragma unroll
      for (int s = 0; s < 3; ++s)
        if (j > 0 && xs[j - 1] > v) { xs[j] = xs[j - 1]; --j; }
      xs[j] = v;
    }
  }
  const long long px = (long long)x << DB_XY_SHIFT;
  // x in [ceil(xa), floor(xb)]  <=>  xa <= x * 2^16  &&  x * 2^16 <= xb - (xb mod 2^16)  <=>  px <= xb
  if (n >= 2 && xs[0] <= px && px <= xs[1]) return true;
  if (n >= 4 && xs[2] <= px && px <= xs[3]) return true;
  return false;
}

// grid (tiles of DB_TH x DB_TW pixels, frames).  The CTA first collects, in descending box order, the boxes whose
// vertex bounding box meets its tile (one thread per box, ballot compaction), then every thread walks that short
// list for its pixel: the first hit is the last box painted.
constexpr int DB_TW = 32, DB_TH = DB_THREADS / DB_TW;
constexpr int DB_LIST = 1024;   // candidates kept in shared memory; a fuller tile falls back to the global walk

__global__ void __launch_bounds__(DB_THREADS) db_raster_kernel(DbParams p, int tiles_x) {
  __shared__ int s_list[DB_LIST];
  __shared__ int s_warp[DB_THREADS / 32];
  __shared__ int s_n;
  const int f = blockIdx.y;
  const int64_t b0 = p.box_off[f], b1 = p.box_off[f + 1];
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int ya = ty * DB_TH, xa = tx * DB_TW;
  const int yb = min(ya + DB_TH, p.H) - 1, xb = min(xa + DB_TW, p.W) - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  bool overflow = false;
  for (int64_t hi = b1; hi > b0; hi -= DB_THREADS) {       // chunks of boxes, last chunk first
    const int64_t b = hi - 1 - threadIdx.x;                // thread 0 looks at the last box of the chunk
    bool keep = false;
    if (b >= b0) {
      const DbBox& r = p.recs[b];
      keep = r.vy_min <= yb && r.vy_max >= ya && r.vx_min <= xb && r.vx_max >= xa;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_n;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (keep && pos < DB_LIST) s_list[pos] = (int)(b - b0);
    __syncthreads();
    if (threadIdx.x == DB_THREADS - 1) s_n = before + __popc(bal);
    __syncthreads();
    if (s_n > DB_LIST) { overflow = true; break; }         // uniform: s_n is shared
  }
  const int y = ya + threadIdx.x / DB_TW, x = xa + (threadIdx.x % DB_TW);
  if (y >= p.H || x >= p.W) return;
  int color = 0;
  if (!overflow) {
    const int n = s_n;
    for (int k = 0; k < n; ++k) {
      const DbBox& r = p.recs[b0 + s_list[k]];
      if (y < r.vy_min || y > r.vy_max) continue;
      bool hit = db_filled(r, x, y);
#pragma unroll
      for (int i = 0; i < 4; ++i) hit = hit || db_on_line(r.line[i], x, y);
      if (hit) {
        color = r.color;
        break;
      }
    }
  } else {
    for (int64_t b = b1 - 1; b >= b0; --b) {
      const DbBox& r = p.recs[b];
      if (y < r.vy_min || y > r.vy_max) continue;
      bool hit = db_filled(r, x, y);
#pragma unroll
      for (int i = 0; i < 4; ++i) hit = hit || db_on_line(r.line[i], x, y);
      if (hit) {
        color = r.color;
        break;
      }
    }
  }
  p.target[((int64_t)f * p.H + y) * p.W + x] = (uint8_t)color;
}

}  // namespace

extern "C" int lv_draw_boxes(lv_handle* h, const double* d_corners, const int32_t* d_colors, int32_t n_frames,
                             const int64_t* h_box_offsets, const int32_t shape[3], const double voxel_size[3],
                             double z_offset, uint8_t* d_target, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_draw_boxes: null handle");
  LV_REQUIRE(shape && voxel_size && h_box_offsets, "lv_draw_boxes: null argument");
  LV_REQUIRE(n_frames >= 0, "lv_draw_boxes: negative frame count");
  LV_REQUIRE(shape[0] > 0 && shape[1] > 0 && shape[0] <= 32768 && shape[1] <= 32768, "lv_draw_boxes: bad image shape %d x %d", shape[0], shape[1]);
  LV_REQUIRE(voxel_size[0] > 0 && voxel_size[1] > 0 && voxel_size[2] > 0, "lv_draw_boxes: voxel_size must be > 0");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_box_offsets[0] == 0, "lv_draw_boxes: box_offsets[0] must be 0");
  for (int f = 0; f < n_frames; ++f)
    LV_REQUIRE(h_box_offsets[f + 1] >= h_box_offsets[f], "lv_draw_boxes: box_offsets must be non-decreasing");
  const int64_t n_boxes = h_box_offsets[n_frames];
  LV_REQUIRE(d_target != nullptr, "lv_draw_boxes: null target");
  LV_REQUIRE(n_boxes == 0 || (d_corners && d_colors), "lv_draw_boxes: null boxes");
  (void)z_offset;  // the z row of the voxel-space matrix does not reach the (x, y) footprint
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK_CUDA(cudaSetDevice(h->device));

  DbParams p;
  memset(&p, 0, sizeof(p));
  p.corners = d_corners; p.colors = d_colors; p.target = d_target;
  // the image is (rows, cols) = (shape[0], shape[1]); x uses shape[0]/2, y uses shape[1]/2
  // (create_transformation_matrix_to_voxel_space, generating_train_bev.py:47-62)
  p.H = shape[0]; p.W = shape[1];
  p.m0 = 1.0 / voxel_size[0]; p.m1 = 1.0 / voxel_size[1];
  p.t0 = (double)shape[0] / 2.0; p.t1 = (double)shape[1] / 2.0;
  p.n_boxes = n_boxes;
  const void* d_off = nullptr;
  LV_CHECK(h->draw_offsets.sync(h_box_offsets, sizeof(int64_t) * (n_frames + 1), stream, &d_off));
  p.box_off = (const int64_t*)d_off;
  LV_CHECK(h->draw_recs.ensure((size_t)(n_boxes + 1) * sizeof(DbBox), stream));
  p.recs = h->draw_recs.as<DbBox>();
  if (n_boxes > 0) {
    db_setup_kernel<<<(unsigned)lv_div_up(n_boxes, DB_THREADS), DB_THREADS, 0, stream>>>(p);
    LV_LAUNCH_CHECK(h);
  }
  const int tiles_x = (int)lv_div_up(p.W, DB_TW), tiles_y = (int)lv_div_up(p.H, DB_TH);
  LV_REQUIRE(n_frames <= 65535, "lv_draw_boxes: at most 65535 frames per call");
  db_raster_kernel<<<dim3((unsigned)(tiles_x * tiles_y), (unsigned)n_frames), DB_THREADS, 0, stream>>>(p, tiles_x);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_draw_boxes_host(lv_handle* h, const double* h_corners, const int32_t* h_colors, int32_t n_frames,
                                  const int64_t* h_box_offsets, const int32_t shape[3], const double voxel_size[3],
                                  double z_offset, uint8_t* h_target) {
  LV_REQUIRE(h != nullptr, "lv_draw_boxes_host: null handle");
  LV_REQUIRE(shape && h_box_offsets && n_frames >= 0, "lv_draw_boxes_host: bad arguments");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_target != nullptr, "lv_draw_boxes_host: null target");
  LV_REQUIRE(shape[0] > 0 && shape[1] > 0, "lv_draw_boxes_host: bad image shape");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const int64_t n_boxes = h_box_offsets[n_frames];
  LV_REQUIRE(n_boxes >= 0 && (n_boxes == 0 || (h_corners && h_colors)), "lv_draw_boxes_host: null boxes");
  const size_t out_bytes = (size_t)n_frames * shape[0] * shape[1];
  LV_CHECK(h->draw_stage[0].ensure((size_t)n_boxes * 12 * sizeof(double), st));
  LV_CHECK(h->draw_stage[1].ensure((size_t)n_boxes * sizeof(int32_t), st));
  LV_CHECK(h->draw_stage[2].ensure(out_bytes, st));
  if (n_boxes > 0) {
    LV_CHECK_CUDA(cudaMemcpyAsync(h->draw_stage[0].ptr, h_corners, (size_t)n_boxes * 12 * sizeof(double), cudaMemcpyHostToDevice, st));
    LV_CHECK_CUDA(cudaMemcpyAsync(h->draw_stage[1].ptr, h_colors, (size_t)n_boxes * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  LV_CHECK(lv_draw_boxes(h, h->draw_stage[0].as<double>(), h->draw_stage[1].as<int32_t>(), n_frames, h_box_offsets, shape,
                         voxel_size, z_offset, h->draw_stage[2].as<uint8_t>(), st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h_target, h->draw_stage[2].ptr, out_bytes, cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  return LV_OK;
}
