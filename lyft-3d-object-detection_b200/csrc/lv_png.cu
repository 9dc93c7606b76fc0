// PNG encoder on the device (sm_100a): the on-disk side of the BEV dataset generator.
//
// Replaces cv2.imwrite(path, bev_im) / cv2.imwrite(path, target_im) of
// generating-dataset/generating_train_bev.py:215, :224, :229 (8-bit colour / grey PNG) for images that
// already live in HBM, so that an encoded ~20 KB file crosses PCIe instead of a 338 KB dense image that is
// 98 % zeros.  PNG is lossless: the contract is decode-exactness under cv2.imread (dataset.py:83-90); the
// byte stream itself is stated by oracle/png_oracle.py and reproduced byte for byte:
//
//   signature | IHDR | IDAT = zlib(78 01 | ONE fixed-Huffman deflate block | Adler-32) | IEND
//   scanline = filter byte 0 + pixels (3-channel images written R,G,B from the array's B,G,R like cv2.imwrite);
//   every scanline is tokenised on its own: a run of R equal bytes = literal + distance-1 matches of
//   min(rest, 258) while rest >= 3 + `rest` literals.
//
// A deflate stream is a BIT string, so a parallel encoder has to know where every piece starts:
//
//   png_rows_kernel    one CTA = 8 scanlines of one frame, one warp per scanline.  The warp loads its row
//                      (coalesced), marks run starts by ballots, counts the row's bits; a decoupled look-back
//                      over the CTAs of the frame (the single-pass scan of vx_assign_kernel) gives the CTA's
//                      bit offset; the warps then emit their tokens into a shared-memory bit buffer whose byte
//                      grid is the stream's.  A CTA owns the bytes that START inside its bit range; the byte
//                      it shares with its successor is completed with the successor's first bits, which are
//                      constant (every scanline starts with the literal of filter byte 0) - no atomics on
//                      global memory, no zero-filled output.  Owned bytes are stored, their raw CRC-32 and
//                      the Adler-32 partial sums of the rows go to a per-CTA record.
//   png_finish_kernel  one warp per frame: combines the records (CRC-32 by GF(2) shifts x^(8n) mod P,
//                      Adler-32 from the weighted sums), writes signature, IHDR, IDAT header, zlib header,
//                      Adler-32, IDAT CRC, IEND and the file size.
//
// `out` may be device memory or MAPPED PINNED HOST memory: in the second case the file bytes go
// straight over PCIe as the kernels produce them and nothing else is transferred (engine.HostPipeline).
// Algorithmic bytes: H*W*ch read + the file size written per frame.
#include "lv_common.cuh"

#define PNG_WARPS 8
#define PNG_THREADS (PNG_WARPS * 32)
#define PNG_HEAD 43            // signature 8 + IHDR chunk 25 + IDAT length 4 + "IDAT" 4 + zlib header 2
#define PNG_TAIL 20            // Adler-32 4 + IDAT CRC 4 + IEND chunk 12
#define PNG_POLY 0xedb88320u

struct PngRec {
  uint32_t crc_raw;            // raw CRC-32 (zero init, no final xor) of the CTA's owned bytes
  uint32_t n_owned;            // owned bytes
  uint32_t s0;                 // sum of the filtered bytes of the CTA's rows
  uint32_t pad;
  unsigned long long s1;       // sum of (index in the frame's filtered stream) * byte
};

struct PngParams {
  const uint8_t* img;          // [F][H][W][ch]
  int H, W, ch, swap_rb;
  int rowbytes;                // 1 + W*ch
  int nwords;                  // ceil(rowbytes / 32)
  int rb_pad;                  // row buffer bytes per warp (multiple of 16)
  int bit_words;               // words of the CTA's bit buffer
  int cpf;                     // CTAs per frame
  uint8_t* out;
  long long out_stride;
  int32_t* sizes;
  unsigned long long* state;   // [F*cpf] look-back descriptors (flag << 62 | bits), zero before the launch
  PngRec* rec;                 // [F*cpf]
  uint32_t x2n[32];            // x^(2^k) mod P (reflected), k = 0..31
  uint8_t head[33];            // signature + IHDR chunk (CRC included), built by the host
};

// ---- GF(2) arithmetic of CRC-32 (reflected polynomial; x^0 is bit 31) - zlib's crc32_combine restated
__host__ __device__ __forceinline__ uint32_t png_multmodp(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (int i = 0; i < 32; ++i) {
    if ((a >> (31 - i)) & 1u) p ^= b;
    b = (b >> 1) ^ (PNG_POLY & (0u - (b & 1u)));
  }
  return p;
}
// x^(8 n) mod P
__device__ __forceinline__ uint32_t png_x8n(const uint32_t* x2n, unsigned long long n) {
  uint32_t p = 1u << 31;
  int k = 3;
  while (n) {
    if (n & 1ull) p = png_multmodp(x2n[k & 31], p);
    n >>= 1;
    ++k;
  }
  return p;
}
__device__ __forceinline__ uint32_t png_crc_byte(uint32_t c, uint32_t b) {
  c ^= b;
#pragma unroll
  for (int i = 0; i < 8; ++i) c = (c >> 1) ^ (PNG_POLY & (0u - (c & 1u)));
  return c;
}

// ---- fixed Huffman code (RFC 1951 3.2.6), values ready for LSB-first packing
__device__ __forceinline__ uint32_t png_rev(uint32_t v, int n) { return __brev(v) >> (32 - n); }
__device__ __forceinline__ int png_lit_bits(int v) { return v < 144 ? 8 : 9; }
__device__ __forceinline__ uint32_t png_lit_code(int v) { return v < 144 ? png_rev(0x30 + v, 8) : png_rev(0x190 + (v - 144), 9); }
// match of `len` (3..258) at distance 1: length symbol + extra bits + the 5-bit distance code 0
__device__ __forceinline__ void png_match(int len, uint32_t& val, int& n) {
  int idx, e, extra;
  if (len == 258) { idx = 28; e = 0; extra = 0; }
  else if (len <= 10) { idx = len - 3; e = 0; extra = 0; }
  else {
    const int t = len - 3;
    e = (31 - __clz(t)) - 2;
    idx = 4 + 4 * e + ((t >> e) & 3);
    extra = t & ((1 << e) - 1);
  }
  const int sym = 257 + idx;
  int cn;
  uint32_t code;
  if (sym < 280) { code = png_rev(sym - 256, 7); cn = 7; }
  else { code = png_rev(0xC0 + (sym - 280), 8); cn = 8; }
  val = code | ((uint32_t)extra << cn);
  n = cn + e + 5;
}
__device__ __forceinline__ int png_match_bits(int len) {
  uint32_t v;
  int n;
  png_match(len, v, n);
  return n;
}
// bits of a run of R bytes of value v (oracle/png_oracle.py::row_tokens)
__device__ __forceinline__ int png_run_bits(int v, int R) {
  const int lit = png_lit_bits(v);
  const int rest = R - 1, full = rest / 258, rem = rest - full * 258;
  return lit + full * 13 + (rem >= 3 ? png_match_bits(rem) : rem * lit);
}
__device__ __forceinline__ void png_put(uint32_t* buf, unsigned pos, uint32_t val, int n) {
  const unsigned w = pos >> 5, sh = pos & 31;
  atomicOr(buf + w, val << sh);
  if (sh + n > 32) atomicOr(buf + w + 1, val >> (32 - sh));
}
__device__ __forceinline__ unsigned png_emit_run(uint32_t* buf, unsigned pos, int v, int R) {
  const int lit = png_lit_bits(v);
  const uint32_t lc = png_lit_code(v);
  png_put(buf, pos, lc, lit);
  pos += lit;
  int rest = R - 1;
  while (rest >= 3) {
    const int take = rest < 258 ? rest : 258;
    uint32_t mv;
    int mn;
    png_match(take, mv, mn);
    png_put(buf, pos, mv, mn);
    pos += mn;
    rest -= take;
  }
  for (; rest > 0; --rest) {
    png_put(buf, pos, lc, lit);
    pos += lit;
  }
  return pos;
}
// position of the next run start after bit `b` of word `w` (or n when there is none)
__device__ __forceinline__ int png_next_start(const uint32_t* mask, int nwords, int w, int b, int n) {
  uint32_t m = b < 31 ? (mask[w] >> (b + 1)) << (b + 1) : 0u;
  while (true) {
    if (m) return w * 32 + __ffs(m) - 1;
    if (++w >= nwords) return n;
    m = mask[w];
  }
}

#define PNG_FLAG_AGG (1ull << 62)
#define PNG_FLAG_PREFIX (2ull << 62)
__device__ __forceinline__ unsigned long long png_ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void png_st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(PNG_THREADS) png_rows_kernel(PngParams p) {
  extern __shared__ __align__(16) unsigned char png_smem[];
  uint8_t* rowbuf = png_smem;                                                   // [8][rb_pad]
  uint32_t* mask = reinterpret_cast<uint32_t*>(rowbuf + PNG_WARPS * p.rb_pad);  // [8][nwords] run starts
  uint32_t* wbits = mask + PNG_WARPS * p.nwords;                                // [8][nwords] bits per word of starts
  uint32_t* bitbuf = wbits + PNG_WARPS * p.nwords;                              // [bit_words]
  __shared__ unsigned s_rowbits[PNG_WARPS];
  __shared__ unsigned long long s_excl;
  __shared__ uint32_t s_crc[PNG_WARPS];
  __shared__ unsigned s_s0[PNG_WARPS];
  __shared__ unsigned long long s_s1[PNG_WARPS];

  const int f = blockIdx.x / p.cpf, c = blockIdx.x - f * p.cpf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = c * PNG_WARPS + warp;
  const int n = p.rowbytes;
  uint8_t* rb = rowbuf + warp * p.rb_pad;
  uint32_t* mk = mask + warp * p.nwords;
  uint32_t* wb = wbits + warp * p.nwords;
  for (int i = threadIdx.x; i < p.bit_words; i += PNG_THREADS) bitbuf[i] = 0;

  // ---- phase 1: the filtered scanline, its Adler sums, its run starts, its bits
  unsigned rowbits = 0;
  if (row < p.H) {
    const uint8_t* src = p.img + ((size_t)f * p.H + row) * (size_t)(n - 1);
    unsigned a0 = 0, a1 = 0;
    if (lane == 0) rb[0] = 0;   // filter type 0
    if (p.ch == 3 && p.swap_rb) {
      for (int x = lane; x < p.W; x += 32) {
        const uint8_t b0 = src[3 * x], b1 = src[3 * x + 1], b2 = src[3 * x + 2];
        rb[1 + 3 * x] = b2; rb[2 + 3 * x] = b1; rb[3 + 3 * x] = b0;
        a0 += (unsigned)b2 + b1 + b0;
        a1 += (unsigned)(1 + 3 * x) * b2 + (unsigned)(2 + 3 * x) * b1 + (unsigned)(3 + 3 * x) * b0;
      }
    } else {
      for (int i = lane; i < n - 1; i += 32) {
        const uint8_t b = src[i];
        rb[1 + i] = b;
        a0 += b;
        a1 += (unsigned)(1 + i) * b;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) {
      s_s0[warp] = a0;
      s_s1[warp] = (unsigned long long)row * n * a0 + a1;
    }
    __syncwarp();
    for (int w = 0; w < p.nwords; ++w) {
      const int i = w * 32 + lane;
      const bool st = i < n && (i == 0 || rb[i] != rb[i - 1]);
      const unsigned m = __ballot_sync(0xffffffffu, st);
      if (lane == 0) mk[w] = m;
    }
    __syncwarp();
    for (int w = lane; w < p.nwords; w += 32) {
      unsigned bits = 0;
      uint32_t m = mk[w];
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int s = w * 32 + b;
        const int e = m ? w * 32 + __ffs(m) - 1 : png_next_start(mk, p.nwords, w, 31, n);
        bits += png_run_bits(rb[s], e - s);
      }
      wb[w] = bits;
      rowbits += bits;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rowbits += __shfl_xor_sync(0xffffffffu, rowbits, o);
  } else if (lane == 0) {
    s_s0[warp] = 0;
    s_s1[warp] = 0;
  }
  if (lane == 0) s_rowbits[warp] = rowbits;
  __syncthreads();

  // ---- bit offset of the CTA inside the frame's deflate stream: look-back over the earlier CTAs
  unsigned woff = (c == 0) ? 3u : 0u;   // the block header BFINAL=1, BTYPE=01 precedes row 0
  unsigned agg = woff;
#pragma unroll
  for (int w = 0; w < PNG_WARPS; ++w) {
    if (w < warp) woff += s_rowbits[w];
    agg += s_rowbits[w];
  }
  const bool last = c == p.cpf - 1;
  if (last) agg += 7;                   // end of block
  if (warp == 0) {
    unsigned long long* st = p.state + blockIdx.x;
    unsigned long long excl = 0;
    if (c == 0) {
      if (lane == 0) png_st_state(st, PNG_FLAG_PREFIX | agg);
    } else {
      if (lane == 0) png_st_state(st, PNG_FLAG_AGG | agg);
      int look = c - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long v = PNG_FLAG_PREFIX;
        if (idx >= 0) {
          do { v = png_ld_state(st - (c - idx)); } while ((v >> 62) == 0);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first_p = __ffs(is_prefix) - 1;
        unsigned long long contrib = (first_p < 0 || lane <= first_p) ? (v & ((1ull << 62) - 1)) : 0ull;
        if (idx < 0) contrib = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (first_p >= 0) break;
        look -= 32;
      }
      if (lane == 0) png_st_state(st, PNG_FLAG_PREFIX | (excl + agg));
    }
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  const unsigned long long excl = s_excl;
  const unsigned sh = (unsigned)(excl & 7);        // buffer bit 0 = stream bit (excl & ~7)

  // ---- phase 2: tokens into the bit buffer
  if (c == 0 && threadIdx.x == 0) png_put(bitbuf, 0, 3u, 3);
  if (row < p.H) {
    // exclusive scan of the per-word bits across the row's words (word w belongs to lane w % 32)
    unsigned carry = sh + woff;
    for (int w0 = 0; w0 < p.nwords; w0 += 32) {
      const int w = w0 + lane;
      const unsigned v = w < p.nwords ? wb[w] : 0;
      unsigned inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      unsigned pos = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
      if (w < p.nwords) {
        uint32_t m = mk[w];
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int s = w * 32 + b;
          const int e = m ? w * 32 + __ffs(m) - 1 : png_next_start(mk, p.nwords, w, 31, n);
          pos = png_emit_run(bitbuf, pos, rb[s], e - s);
        }
      }
    }
  }
  // the byte shared with the next CTA is completed with that CTA's first bits: the literal of filter byte 0
  const unsigned endbit = sh + agg;
  if (!last && (endbit & 7) && threadIdx.x == 32) png_put(bitbuf, endbit, 0x0Cu, 8);
  __syncthreads();

  // ---- owned bytes [j0, j1) of the buffer = stream bytes [excl/8 + j0, ...): store + raw CRC-32
  const unsigned j0 = sh ? 1u : 0u, j1 = (endbit + 7) >> 3;
  const unsigned n_owned = j1 - j0;
  const long long k0 = (long long)(excl >> 3) + j0;
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(bitbuf);
  uint8_t* dst = p.out + (size_t)f * p.out_stride + PNG_HEAD + k0;
  const long long room = p.out_stride - PNG_HEAD - PNG_TAIL - k0;   // bytes this CTA may write
  for (unsigned i = threadIdx.x; i < n_owned; i += PNG_THREADS)
    if ((long long)i < room) dst[i] = bytes[j0 + i];
  // slices of equal length, the ragged one in FRONT (leading zero bytes do not change a raw CRC)
  const unsigned slice = (n_owned + PNG_THREADS - 1) / PNG_THREADS;
  const unsigned pad = slice * PNG_THREADS - n_owned;
  uint32_t crc = 0;
  for (unsigned q = 0; q < slice; ++q) {
    const unsigned idx = threadIdx.x * slice + q;
    if (idx >= pad) crc = png_crc_byte(crc, bytes[j0 + idx - pad]);
  }
  uint32_t xp = png_x8n(p.x2n, slice);   // x^(8 * slice)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t right = __shfl_down_sync(0xffffffffu, crc, d);
    crc = png_multmodp(xp, crc) ^ right;   // valid in the lanes that are multiples of 2d
    xp = png_multmodp(xp, xp);
  }
  if (lane == 0) s_crc[warp] = crc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    unsigned s0 = 0;
    unsigned long long s1 = 0;
    for (int w = 0; w < PNG_WARPS; ++w) {
      acc = png_multmodp(xp, acc) ^ s_crc[w];   // xp = x^(8 * 32 * slice)
      s0 += s_s0[w];
      s1 += s_s1[w];
    }
    PngRec r;
    r.crc_raw = acc; r.n_owned = n_owned; r.s0 = s0; r.pad = 0; r.s1 = s1;
    p.rec[blockIdx.x] = r;
  }
}

__device__ __forceinline__ void png_store_be32(uint8_t* d, uint32_t v) {
  d[0] = (uint8_t)(v >> 24); d[1] = (uint8_t)(v >> 16); d[2] = (uint8_t)(v >> 8); d[3] = (uint8_t)v;
}

__global__ void __launch_bounds__(PNG_THREADS) png_finish_kernel(PngParams p, int n_frames) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * PNG_WARPS + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const PngRec* rec = p.rec + (size_t)f * p.cpf;
  // total bytes of the deflate stream, Adler sums
  unsigned long long D = 0, S0 = 0, S1 = 0;
  for (int c = lane; c < p.cpf; c += 32) {
    D += rec[c].n_owned;
    S0 += rec[c].s0;
    S1 += rec[c].s1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    D += __shfl_xor_sync(0xffffffffu, D, o);
    S0 += __shfl_xor_sync(0xffffffffu, S0, o);
    S1 += __shfl_xor_sync(0xffffffffu, S1, o);
  }
  const unsigned long long nfil = (unsigned long long)p.H * p.rowbytes;
  const uint32_t A = (uint32_t)((1ull + S0) % 65521ull);
  const uint32_t B = (uint32_t)((nfil + nfil * S0 - S1) % 65521ull);   // sum over i of (n - i) d_i, plus n * 1
  const uint32_t adler = (B << 16) | A;
  // CRC-32 of the IDAT chunk: "IDAT" 78 01 | deflate bytes | Adler-32
  uint32_t part = 0;
  unsigned long long done = 0;   // owned bytes of the CTAs before this round
  for (int c0 = 0; c0 < p.cpf; c0 += 32) {
    const int c = c0 + lane;
    const unsigned long long mine = c < p.cpf ? rec[c].n_owned : 0;
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (c < p.cpf) part ^= png_multmodp(png_x8n(p.x2n, D - (done + inc) + 4), rec[c].crc_raw);
    done += __shfl_sync(0xffffffffu, inc, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
  if (lane != 0) return;
  const uint8_t pre[6] = {'I', 'D', 'A', 'T', 0x78, 0x01};
  uint32_t c_pre = 0, c_ad = 0;
  for (int i = 0; i < 6; ++i) c_pre = png_crc_byte(c_pre, pre[i]);
  for (int i = 0; i < 4; ++i) c_ad = png_crc_byte(c_ad, (adler >> (24 - 8 * i)) & 0xffu);
  const unsigned long long M = 6 + D + 4;
  uint32_t raw = png_multmodp(png_x8n(p.x2n, D + 4), c_pre) ^ part ^ c_ad;
  const uint32_t crc = raw ^ png_multmodp(png_x8n(p.x2n, M), 0xffffffffu) ^ 0xffffffffu;
  const long long total = PNG_HEAD + (long long)D + PNG_TAIL;
  if (total > p.out_stride) {
    p.sizes[f] = (int32_t)(-total);   // the slot is too small: nothing usable was written
    return;
  }
  uint8_t* o = p.out + (size_t)f * p.out_stride;
  for (int i = 0; i < 33; ++i) o[i] = p.head[i];
  png_store_be32(o + 33, (uint32_t)(2 + D + 4));
  for (int i = 0; i < 6; ++i) o[37 + i] = pre[i];
  uint8_t* t = o + PNG_HEAD + D;
  png_store_be32(t, adler);
  png_store_be32(t + 4, crc);
  const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
  for (int i = 0; i < 12; ++i) t[8 + i] = iend[i];
  p.sizes[f] = (int32_t)total;
}

static uint32_t png_host_crc(const uint8_t* d, size_t n) {
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (PNG_POLY & (0u - (c & 1u)));
  }
  return c ^ 0xffffffffu;
}

extern "C" int64_t lv_png_max_bytes(int32_t height, int32_t width, int32_t channels) {
  if (height <= 0 || width <= 0 || (channels != 1 && channels != 3)) return LV_E_INVALID;
  // every byte a 9-bit literal, header and end-of-block bits
  const int64_t bits = (int64_t)height * (1 + (int64_t)width * channels) * 9 + 10;
  return PNG_HEAD + (bits + 7) / 8 + PNG_TAIL;
}

extern "C" int lv_png_encode(lv_handle* h, const uint8_t* d_images, int32_t n_frames, int32_t height, int32_t width,
                             int32_t channels, int32_t swap_rb, uint8_t* out, int64_t out_stride, int32_t* sizes,
                             lv_stream stream_) {
  LV_REQUIRE(h != nullptr, "lv_png_encode: null handle");
  LV_REQUIRE(n_frames >= 0 && height > 0 && width > 0, "lv_png_encode: bad sizes");
  LV_REQUIRE(channels == 1 || channels == 3, "lv_png_encode: 1 (grey) or 3 (colour) channels, got %d", channels);
  LV_REQUIRE((int64_t)width * channels + 1 <= 4096, "lv_png_encode: scanlines of at most 4096 bytes (width %d x %d channels)",
             width, channels);
  LV_REQUIRE(out_stride >= PNG_HEAD + PNG_TAIL + 8 && out_stride < (1ll << 31), "lv_png_encode: bad out_stride %lld",
             (long long)out_stride);
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(d_images && out && sizes, "lv_png_encode: null pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t stream = (cudaStream_t)stream_;
  PngParams p;
  memset(&p, 0, sizeof(p));
  p.img = d_images; p.H = height; p.W = width; p.ch = channels; p.swap_rb = swap_rb ? 1 : 0;
  p.rowbytes = 1 + width * channels;
  p.nwords = (p.rowbytes + 31) / 32;
  p.rb_pad = (int)lv_align_up((size_t)p.rowbytes + 1, 16);
  p.bit_words = (PNG_WARPS * p.rowbytes * 9 + 10 + 7 + 8 + 31) / 32 + 2;
  p.cpf = (int)lv_div_up(height, PNG_WARPS);
  p.out = out; p.out_stride = out_stride; p.sizes = sizes;
  const int64_t n_cta = (int64_t)n_frames * p.cpf;
  LV_REQUIRE(n_cta < (1ll << 31), "lv_png_encode: too many scanlines");
  LV_CHECK(h->png_state.ensure((size_t)n_cta * sizeof(unsigned long long), stream));
  LV_CHECK(h->png_rec.ensure((size_t)n_cta * sizeof(PngRec), stream));
  p.state = h->png_state.as<unsigned long long>();
  p.rec = h->png_rec.as<PngRec>();
  p.x2n[0] = 1u << 30;   // x^1
  for (int k = 1; k < 32; ++k) p.x2n[k] = png_multmodp(p.x2n[k - 1], p.x2n[k - 1]);
  const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memcpy(p.head, sig, 8);
  uint8_t* ih = p.head + 8;
  ih[0] = 0; ih[1] = 0; ih[2] = 0; ih[3] = 13;
  memcpy(ih + 4, "IHDR", 4);
  const uint32_t wh[2] = {(uint32_t)width, (uint32_t)height};
  for (int k = 0; k < 2; ++k)
    for (int b = 0; b < 4; ++b) ih[8 + 4 * k + b] = (uint8_t)(wh[k] >> (24 - 8 * b));
  ih[16] = 8; ih[17] = channels == 3 ? 2 : 0; ih[18] = 0; ih[19] = 0; ih[20] = 0;
  const uint32_t hc = png_host_crc(ih + 4, 17);
  for (int b = 0; b < 4; ++b) ih[21 + b] = (uint8_t)(hc >> (24 - 8 * b));
  const size_t smem = (size_t)PNG_WARPS * p.rb_pad + (size_t)2 * PNG_WARPS * p.nwords * 4 + (size_t)p.bit_words * 4;
  if (smem > 40 * 1024)
    LV_CHECK_CUDA(cudaFuncSetAttribute(png_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LV_CHECK_CUDA(cudaMemsetAsync(p.state, 0, (size_t)n_cta * sizeof(unsigned long long), stream));
  png_rows_kernel<<<(unsigned)n_cta, PNG_THREADS, smem, stream>>>(p);
  LV_LAUNCH_CHECK(h);
  png_finish_kernel<<<(unsigned)lv_div_up(n_frames, PNG_WARPS), PNG_THREADS, 0, stream>>>(p, n_frames);
  LV_LAUNCH_CHECK(h);
  return LV_OK;
}

// Host-buffer form (the cv2.imwrite replacement of a caller that holds numpy images): images in, files out.
extern "C" int lv_png_encode_host(lv_handle* h, const uint8_t* h_images, int32_t n_frames, int32_t height, int32_t width,
                                  int32_t channels, int32_t swap_rb, uint8_t* h_out, int64_t out_stride, int32_t* h_sizes) {
  LV_REQUIRE(h != nullptr, "lv_png_encode_host: null handle");
  LV_REQUIRE(n_frames >= 0 && height > 0 && width > 0 && (channels == 1 || channels == 3), "lv_png_encode_host: bad sizes");
  if (n_frames == 0) return LV_OK;
  LV_REQUIRE(h_images && h_out && h_sizes, "lv_png_encode_host: null pointer");
  LV_CHECK_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  const size_t in_bytes = (size_t)n_frames * height * width * channels;
  LV_CHECK(h->png_stage[0].ensure(in_bytes, st));
  LV_CHECK(h->png_stage[1].ensure((size_t)n_frames * out_stride, st));
  LV_CHECK(h->png_stage[2].ensure((size_t)n_frames * 4, st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h->png_stage[0].ptr, h_images, in_bytes, cudaMemcpyHostToDevice, st));
  LV_CHECK(lv_png_encode(h, h->png_stage[0].as<uint8_t>(), n_frames, height, width, channels, swap_rb,
                         h->png_stage[1].as<uint8_t>(), out_stride, h->png_stage[2].as<int32_t>(), st));
  LV_CHECK_CUDA(cudaMemcpyAsync(h_sizes, h->png_stage[2].ptr, (size_t)n_frames * 4, cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  for (int32_t f = 0; f < n_frames; ++f)
    if (h_sizes[f] > 0)
      LV_CHECK_CUDA(cudaMemcpyAsync(h_out + (size_t)f * out_stride, h->png_stage[1].as<uint8_t>() + (size_t)f * out_stride,
                                    (size_t)h_sizes[f], cudaMemcpyDeviceToHost, st));
  LV_CHECK_CUDA(cudaStreamSynchronize(st));
  return LV_OK;
}
