// Half-precision forms of the PointPillars decoration and scatter (sm_100a).
//
// The reference trains its nuScenes/Lyft PointPillars configs under apex O2 (`enable_mixed_precision: true`,
// second/configs/nuscenes/all.pp.mida.config:348): second/second/pytorch/train.py:34-47 casts "voxels" to half
// before the network sees them, so PillarFeatureNet*.forward (pointpillars.py:203-231 and variants) and
// PointPillarsScatter.forward (:444-476) run on float16 tensors.
//
// pillar_decorate_h_kernel  every torch op on half tensors computes in float32 and rounds its RESULT to half; the
//   kernel reproduces those roundings statement by statement (oracle/pillar_oracle.py::decorate_half, checked bit
//   for bit against the reference's own classes run in half):  mean = rh(rh(sum x) / num), f_cluster = rh(x - mean),
//   centre = rh(rh(coor * vx) + x_offset), f_center = rh(x - centre), radius / distance = rh(sqrt(sum of squares)).
//   One warp per pillar, the pillar (T <= 64 points of 4 halves = 8 bytes each) lives in registers, the (T, C_out)
//   half row leaves through a shared-memory stage as 32-bit stores.
//   Algorithmic bytes per pillar: T*8 + 20 read, T*C_out*2 written.
// pillar_canvas_h_kernel    the scatter-as-gather of lv_pillar.cu writing a HALF canvas in one pass (2.6 GB
//   instead of 5.2 GB for 128 samples of 64 x 400 x 400): one CTA = 256 consecutive cells x all channels, every lane
//   owns 8 cells (one 128-bit store per channel row), a warp owns C/8 contiguous channels so that a cell's features
//   arrive as 128-bit loads of 8 channels, transposed in registers by byte permutes.
//   Algorithmic bytes: P*(C*2+16) read, B*C*ny*nx*2 written.
#include <cuda_fp16.h>

#include "lv_common.cuh"

#define PH_WARPS 8

__device__ __forceinline__ float ph_rh(float v) { return __half2float(__float2half_rn(v)); }
__device__ __forceinline__ float ph_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ph_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float ph_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct HalfDecoParams {
  const uint2* voxels;   // (P, T) points of 4 halves
  const int32_t* num;
  const int4* coors;     // (P) b, z, y, x
  int64_t P;
  int T, C_out, variant, with_distance;
  float vx, vy, x_off, y_off;
  __half* out;           // (P, T, C_out)
};

struct HPoint { float x, y, z, w; };
__device__ __forceinline__ HPoint ph_unpack(uint2 q) {
  const __half2 a = *reinterpret_cast<const __half2*>(&q.x), b = *reinterpret_cast<const __half2*>(&q.y);
  HPoint r;
  r.x = __low2float(a); r.y = __high2float(a); r.z = __low2float(b); r.w = __high2float(b);
  return r;
}

// one live slot -> C_out halves at o
__device__ __forceinline__ void ph_slot(const HPoint q, float mx, float my, float mz, float cx, float cy, float height,
                                        int variant, int with_distance, __half* o) {
  const float px = ph_rh(q.x - cx), py = ph_rh(q.y - cy);                       // f_center (:213-217)
  int k;
  if (variant == LV_PILLAR_PFN) {
    o[0] = __float2half_rn(q.x); o[1] = __float2half_rn(q.y); o[2] = __float2half_rn(q.z); o[3] = __float2half_rn(q.w);
    k = 4;
  } else if (variant == LV_PILLAR_OLD) {
    o[0] = __float2half_rn(px); o[1] = __float2half_rn(py); o[2] = __float2half_rn(q.z); o[3] = __float2half_rn(q.w);
    k = 4;
  } else {                                                                      // :303-306
    o[0] = __float2half_rn(sqrtf(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y))));
    o[1] = __float2half_rn(q.z); o[2] = __float2half_rn(q.w);
    k = 3;
  }
  o[k] = __float2half_rn(q.x - mx); o[k + 1] = __float2half_rn(q.y - my); o[k + 2] = __float2half_rn(q.z - mz);   // f_cluster (:210)
  o[k + 3] = __float2half_rn(px); o[k + 4] = __float2half_rn(py);
  k += 5;
  if (variant == LV_PILLAR_RADIUS_HEIGHT) o[k++] = __float2half_rn(height);
  if (with_distance) {
    const float dx = variant == LV_PILLAR_OLD ? px : q.x, dy = variant == LV_PILLAR_OLD ? py : q.y;
    o[k] = __float2half_rn(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(q.z, q.z))));
  }
}

__global__ void __launch_bounds__(PH_WARPS * 32) pillar_decorate_h_kernel(HalfDecoParams p) {
  extern __shared__ __align__(16) __half hstage[];   // [PH_WARPS][T*C_out + 2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = p.T * p.C_out;
  __half* st = hstage + warp * ((per + 2 + 1) & ~1);
  const int64_t warps_total = (int64_t)gridDim.x * PH_WARPS;
  for (int64_t pil = (int64_t)blockIdx.x * PH_WARPS + warp; pil < p.P; pil += warps_total) {
    const uint2* v = p.voxels + pil * p.T;
    HPoint a = {0.f, 0.f, 0.f, 0.f}, b = a;
    if (lane < p.T) a = ph_unpack(__ldg(v + lane));
    if (lane + 32 < p.T) b = ph_unpack(__ldg(v + lane + 32));
    int num = __ldg(p.num + pil);
    const int4 co = __ldg(p.coors + pil);
    // mean over ALL T slots, padding zeros included (:208-209): float32 accumulation, one rounding for the sum,
    // one for the quotient
    const float fn = ph_rh((float)num);
    const float mx = ph_rh(__fdiv_rn(ph_rh(ph_sum(a.x + b.x)), fn));
    const float my = ph_rh(__fdiv_rn(ph_rh(ph_sum(a.y + b.y)), fn));
    const float mz = ph_rh(__fdiv_rn(ph_rh(ph_sum(a.z + b.z)), fn));
    float height = 0.f;
    if (p.variant == LV_PILLAR_RADIUS_HEIGHT) {   // :387-389, min/max over all T slots
      const bool ha = lane < p.T, hb = lane + 32 < p.T;
      const float zmax = fmaxf(ha ? a.z : -INFINITY, hb ? b.z : -INFINITY);
      const float zmin = fminf(ha ? a.z : INFINITY, hb ? b.z : INFINITY);
      height = ph_rh(ph_max(zmax) - ph_min(zmin));
    }
    const float cx = ph_rh(__fadd_rn(ph_rh(__fmul_rn((float)co.w, p.vx)), p.x_off));
    const float cy = ph_rh(__fadd_rn(ph_rh(__fmul_rn((float)co.z, p.vy)), p.y_off));
    num = num < 0 ? 0 : (num > p.T ? p.T : num);
    if (lane < num) ph_slot(a, mx, my, mz, cx, cy, height, p.variant, p.with_distance, st + lane * p.C_out);
    if (lane + 32 < num) ph_slot(b, mx, my, mz, cx, cy, height, p.variant, p.with_distance, st + (lane + 32) * p.C_out);
    const int nd = num * p.C_out;
    if (lane == 0 && (nd & 1)) st[nd] = __float2half_rn(0.f);   // pad the last pair of the data part
    __syncwarp();
    __half* dst = p.out + pil * per;
    if ((per & 1) == 0) {   // rows start 4-byte aligned: pairs of halves, zeros from registers beyond the data
      const uint32_t* s2 = reinterpret_cast<const uint32_t*>(st);
      uint32_t* d2 = reinterpret_cast<uint32_t*>(dst);
      const int nd2 = (nd + 1) >> 1;
      for (int i = lane; i < (per >> 1); i += 32) d2[i] = i < nd2 ? s2[i] : 0u;
    } else {
      for (int i = lane; i < per; i += 32) dst[i] = i < nd ? st[i] : __float2half_rn(0.f);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------- scatter into a half canvas
#define SH_THREADS 256
#define SH_WARPS 8
#define SH_TILE 256   // cells per CTA: every lane owns 8 consecutive cells (one 128-bit store per channel row)

// half r (0..7) of the eight cells' feature groups -> one 128-bit row store
template <int R>
__device__ __forceinline__ uint4 sh_row(const uint4 (&v)[8]) {
  constexpr unsigned sel = (R & 1) ? 0x7632u : 0x5410u;   // (lo|hi halves of a, b) -> a.r | b.r << 16
  auto word = [](const uint4& q) { return (R >> 1) == 0 ? q.x : (R >> 1) == 1 ? q.y : (R >> 1) == 2 ? q.z : q.w; };
  uint4 o;
  o.x = __byte_perm(word(v[0]), word(v[1]), sel);
  o.y = __byte_perm(word(v[2]), word(v[3]), sel);
  o.z = __byte_perm(word(v[4]), word(v[5]), sel);
  o.w = __byte_perm(word(v[6]), word(v[7]), sel);
  return o;
}
__device__ __forceinline__ void sh_st(__half* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(SH_THREADS, 5) pillar_canvas_h_kernel(const __half* __restrict__ feats, int32_t* __restrict__ map,
                                                                       int C, int64_t ncell, int tiles_per_sample,
                                                                       __half* __restrict__ canvas) {
  const int b = blockIdx.x / tiles_per_sample;
  const int t = blockIdx.x - b * tiles_per_sample;
  const int64_t cell0 = (int64_t)t * SH_TILE;
  int32_t* m = map + (int64_t)b * ncell + cell0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = lane * 8;
  const int cpw = C / SH_WARPS;   // channels per warp, a multiple of 8
  __half* row = canvas + ((int64_t)b * C + warp * cpw) * ncell + cell0 + j0;
  const int4 o0 = *reinterpret_cast<const int4*>(m + j0), o1 = *reinterpret_cast<const int4*>(m + j0 + 4);
  const int occ[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
  const bool mine = (o0.x & o0.y & o0.z & o0.w & o1.x & o1.y & o1.z & o1.w) >= 0;   // any of the eight >= 0
  const bool any = __syncthreads_or(mine);
  if (any && warp == 0 && mine) {   // touched-cell reset: the map is all -1 again when the kernel ends
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (occ[k] >= 0) m[j0 + k] = -1;
  }
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  if (!any) {
    for (int r = 0; r < cpw; ++r, row += ncell) sh_st(row, z);
    return;
  }
  const uint4* f4 = reinterpret_cast<const uint4*>(feats) + ((warp * cpw) >> 3);
  const unsigned c8 = (unsigned)C >> 3;
  for (int g = 0; g < (cpw >> 3); ++g) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = occ[k] >= 0 ? __ldg(f4 + (unsigned)occ[k] * c8 + g) : z;
    sh_st(row, sh_row<0>(v));
    sh_st(row + ncell, sh_row<1>(v));
    sh_st(row + 2 * ncell, sh_row<2>(v));
    sh_st(row + 3 * ncell, sh_row<3>(v));
    sh_st(row + 4 * ncell, sh_row<4>(v));
    sh_st(row + 5 * ncell, sh_row<5>(v));
    sh_st(row + 6 * ncell, sh_row<6>(v));
    sh_st(row + 7 * ncell, sh_row<7>(v));
    row += 8 * ncell;
  }
}

// any shape: one thread per canvas element
__global__ void __launch_bounds__(256) pillar_canvas_h_generic_kernel(const __half* __restrict__ feats, const int32_t* __restrict__ map,
                                                                     int C, int64_t ncell, int64_t total, __half* __restrict__ canvas) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cell = i % ncell;
    const int64_t bc = i / ncell;
    const int c = (int)(bc % C);
    const int64_t b = bc / C;
    const int pil = map[b * ncell + cell];
    canvas[i] = pil >= 0 ? feats[(int64_t)pil * C + c] : __float2half_rn(0.f);
  }
}
__global__ void __launch_bounds__(256) pillar_map_reset_kernel(const int32_t* __restrict__ coords, int64_t P, int B, int ny, int nx,
                                                              int32_t* __restrict__ map) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + p);
  if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  map[(int64_t)c.x * ny * nx + (int64_t)c.z * nx + c.w] = -1;
}

__global__ void __launch_bounds__(256) pillar_index_h_kernel(const int32_t* __restrict__ coords, int64_t P, int B, int ny, int nx,
                                                            int32_t* __restrict__ map) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + p);  // b, z, y, x
  if (c.x < 0 || c.x >= B || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  map[(int64_t)c.x * ny * nx + (int64_t)c.z * nx + c.w] = (int32_t)p;
}

extern "C" int lv_pillar_decorate_half(lv_handle* h, const uint16_t* d_voxels, const int32_t* d_num_points,
                                       const int32_t* d_coors, int64_t n_pillars, int32_t max_points, int32_t num_features,
                                       float vx, float vy, float x_offset, float y_offset, int32_t variant,
                                       int32_t with_distance, uint16_t* d_out, lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_decorate_half: null handle");
  LV_REQUIRE(n_pillars >= 0 && max_points > 0, "lv_pillar_decorate_half: bad sizes");
  LV_REQUIRE(num_features == 4 && max_points <= 64, "lv_pillar_decorate_half: needs 4 features per point and max_points <= 64 "
             "(got %d, %d)", num_features, max_points);
  const int c_out = lv_pillar_out_channels(num_features, variant, with_distance);
  LV_REQUIRE(c_out > 0, "lv_pillar_decorate_half: bad variant %d", variant);
  if (n_pillars == 0) return LV_OK;
  LV_REQUIRE(d_voxels && d_num_points && d_coors && d_out, "lv_pillar_decorate_half: null pointer");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coors) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_voxels) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(d_out) & 3) == 0, "lv_pillar_decorate_half: misaligned pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  HalfDecoParams p;
  p.voxels = reinterpret_cast<const uint2*>(d_voxels);
  p.num = d_num_points;
  p.coors = reinterpret_cast<const int4*>(d_coors);
  p.P = n_pillars; p.T = max_points; p.C_out = c_out; p.variant = variant; p.with_distance = with_distance ? 1 : 0;
  p.vx = vx; p.vy = vy; p.x_off = x_offset; p.y_off = y_offset;
  p.out = reinterpret_cast<__half*>(d_out);
  const size_t smem = (size_t)PH_WARPS * ((max_points * c_out + 3) & ~1) * sizeof(__half);
  int64_t blocks = (int64_t)h->num_sms * 8;
  const int64_t need = lv_div_up(n_pillars, PH_WARPS);
  if (blocks > need) blocks = need;
  pillar_decorate_h_kernel<<<(unsigned)blocks, PH_WARPS * 32, smem, (cudaStream_t)stream_>>>(p);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

extern "C" int lv_pillar_scatter_half(lv_handle* h, const uint16_t* d_feats, const int32_t* d_coords, int64_t n_pillars,
                                      int32_t channels, int32_t batch_size, int32_t ny, int32_t nx, uint16_t* d_canvas,
                                      lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_pillar_scatter_half: null handle");
  LV_REQUIRE(n_pillars >= 0 && channels > 0 && batch_size >= 0 && ny > 0 && nx > 0, "lv_pillar_scatter_half: bad sizes");
  if (batch_size == 0) return LV_OK;
  LV_REQUIRE(d_canvas != nullptr, "lv_pillar_scatter_half: null canvas");
  LV_REQUIRE(n_pillars == 0 || (d_feats && d_coords), "lv_pillar_scatter_half: null pointer");
  LV_REQUIRE(n_pillars < (1ll << 31), "lv_pillar_scatter_half: too many pillars");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(d_coords) & 15) == 0, "lv_pillar_scatter_half: coords must be 16-byte aligned");
  const int64_t ncell = (int64_t)ny * nx;
  LV_REQUIRE(n_pillars * (int64_t)channels < (1ll << 32), "lv_pillar_scatter_half: %lld pillars x %d channels exceed 2^32 features",
             (long long)n_pillars, channels);
  const int tiles = (int)lv_div_up(ncell, SH_TILE);
  const int64_t grid = (int64_t)tiles * batch_size;
  LV_REQUIRE(grid < (1ll << 31), "lv_pillar_scatter_half: canvas too large");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  LV_CHECK(h->pil_map.ensure((size_t)batch_size * ncell * sizeof(int32_t), stream, 0xff));
  int32_t* map = h->pil_map.as<int32_t>();
  if (n_pillars > 0) {
    pillar_index_h_kernel<<<(unsigned)lv_div_up(n_pillars, 256), 256, 0, stream>>>(d_coords, n_pillars, batch_size, ny, nx, map);
    LV_LAUNCH_CHECK(h);
  }
  const __half* feats = reinterpret_cast<const __half*>(d_feats);
  __half* canvas = reinterpret_cast<__half*>(d_canvas);
  const bool fast = ncell % SH_TILE == 0 && channels % 64 == 0 && (reinterpret_cast<uintptr_t>(d_canvas) & 15) == 0 &&
                    (n_pillars == 0 || (reinterpret_cast<uintptr_t>(d_feats) & 15) == 0);
  if (fast) {
    pillar_canvas_h_kernel<<<(unsigned)grid, SH_THREADS, 0, stream>>>(feats, map, channels, ncell, tiles, canvas);
    LV_LAUNCH_CHECK(h);
  } else {
    const int64_t total = (int64_t)batch_size * channels * ncell;
    int64_t blocks = lv_div_up(total, 256);
    if (blocks > (int64_t)h->num_sms * 16) blocks = (int64_t)h->num_sms * 16;
    pillar_canvas_h_generic_kernel<<<(unsigned)blocks, 256, 0, stream>>>(feats, map, channels, ncell, total, canvas);
    LV_LAUNCH_CHECK(h);
    if (n_pillars > 0) {
      pillar_map_reset_kernel<<<(unsigned)lv_div_up(n_pillars, 256), 256, 0, stream>>>(d_coords, n_pillars, batch_size, ny, nx, map);
      LV_LAUNCH_CHECK(h);
    }
  }
  return LV_OK;
}
