// N1 -- TIFF LZW strip / tile decoder for the GeoTIFF reader.
//
// The reference reads its rasters through rasterio / GDAL (TreeDetection/prediction.py:61,
// postprocessing.py:781-800, merging.py:56-75); orthophoto and nDSM deliveries are commonly LZW
// compressed.  This is host code (no kernel): every strip / tile is an independent LZW stream, the
// Python reader decodes them on a thread pool (ctypes releases the GIL) straight into the pinned
// buffer the host->device copy starts from.  Algorithm: TIFF 6.0 section 13 (MSB-first codes of
// 9..12 bits, ClearCode 256, EOI 257, "early change" as libtiff writes it).  Oracle: PIL / libtiff
// decoding the same file (tests/test_geotiff_codec.py).
#include <cstdint>

#include "common.cuh"

// returns the number of bytes written to dst (<= cap), or a negative TD_ERR_* code
extern "C" long long td_tiff_lzw_decode(const unsigned char* src, long long n_src, unsigned char* dst,
                                        long long cap) {
  if (!src || !dst || n_src < 0 || cap < 0) { td_set_error("td_tiff_lzw_decode: bad argument"); return TD_ERR_ARG; }
  static const int kMax = 4096;
  uint16_t prefix[kMax], length[kMax];
  unsigned char suffix[kMax], first[kMax];
  for (int c = 0; c < 256; ++c) { prefix[c] = 0; length[c] = 1; suffix[c] = (unsigned char)c; first[c] = (unsigned char)c; }
  int width = 9, next = 258, prev = -1;
  uint32_t bitbuf = 0;
  int bits = 0;
  long long ip = 0, out = 0;
  for (;;) {
    while (bits < width) {
      if (ip >= n_src) return out;               // stream ended without EOI (tolerated, as libtiff does)
      bitbuf = (bitbuf << 8) | src[ip++];
      bits += 8;
    }
    const int code = (int)((bitbuf >> (bits - width)) & ((1u << width) - 1u));
    bits -= width;
    if (code == 257) break;
    if (code == 256) { width = 9; next = 258; prev = -1; continue; }
    if (prev < 0) {
      if (code >= 256) { td_set_error("td_tiff_lzw_decode: corrupt stream"); return TD_ERR_ARG; }
      if (out + 1 > cap) { td_set_error("td_tiff_lzw_decode: output overflow"); return TD_ERR_OVERFLOW; }
      dst[out++] = (unsigned char)code;
      prev = code;
      continue;
    }
    int len;
    unsigned char fc;
    if (code < next) {
      len = length[code];
      if (out + len > cap) { td_set_error("td_tiff_lzw_decode: output overflow"); return TD_ERR_OVERFLOW; }
      int p = code;
      for (int k = len - 1; k >= 0; --k) { dst[out + k] = suffix[p]; p = prefix[p]; }
      fc = first[code];
    } else if (code == next) {                    // KwKwK: string(prev) + first(prev)
      len = length[prev] + 1;
      if (out + len > cap) { td_set_error("td_tiff_lzw_decode: output overflow"); return TD_ERR_OVERFLOW; }
      int p = prev;
      for (int k = len - 2; k >= 0; --k) { dst[out + k] = suffix[p]; p = prefix[p]; }
      fc = first[prev];
      dst[out + len - 1] = fc;
    } else {
      td_set_error("td_tiff_lzw_decode: corrupt stream");
      return TD_ERR_ARG;
    }
    out += len;
    if (next < kMax) {
      prefix[next] = (uint16_t)prev;
      suffix[next] = fc;
      first[next] = first[prev];
      length[next] = (uint16_t)(length[prev] + 1);
      ++next;
      if (next >= (1 << width) - 1 && width < 12) ++width;
    }
    prev = code;
  }
  return out;
}

// The encoder of the same format (TIFF 6.0 section 13 as libtiff writes it: ClearCode first, MSB-first codes,
// "early change", ClearCode again when the table is full, EOI last), so that this package can WRITE the
// compressed rasters its readers are measured on.  Returns the bytes written or a negative TD_ERR_* code;
// cap >= n_src * 3 / 2 + 16 always suffices.
extern "C" long long td_tiff_lzw_encode(const unsigned char* src, long long n_src, unsigned char* dst, long long cap) {
  if (!src || !dst || n_src < 0 || cap < 0) { td_set_error("td_tiff_lzw_encode: bad argument"); return TD_ERR_ARG; }
  constexpr int kHash = 1 << 14;                       // open addressing, <= 3837 live entries
  static thread_local int32_t hkey[kHash];
  static thread_local uint16_t hval[kHash];
  uint64_t acc = 0;
  int nacc = 0;
  long long out = 0;
  int width = 9, next = 258;
  bool overflow = false;
  auto put = [&](int code) {
    acc = (acc << width) | (uint64_t)code;
    nacc += width;
    while (nacc >= 8) {
      if (out < cap) dst[out] = (unsigned char)(acc >> (nacc - 8)); else overflow = true;
      ++out;
      nacc -= 8;
    }
  };
  auto reset = [&] { for (int i = 0; i < kHash; ++i) hkey[i] = -1; width = 9; next = 258; };
  reset();
  put(256);
  if (n_src > 0) {
    int omega = src[0];
    for (long long i = 1; i < n_src; ++i) {
      const int k = src[i];
      const int32_t key = (omega << 8) | k;
      uint32_t h = ((uint32_t)key * 2654435761u) >> 18;
      bool found = false;
      while (hkey[h] >= 0) {
        if (hkey[h] == key) { found = true; break; }
        h = (h + 1) & (kHash - 1);
      }
      if (found) { omega = hval[h]; continue; }
      put(omega);
      hkey[h] = key;
      hval[h] = (uint16_t)next;
      ++next;
      if (next == 4094) { put(256); reset(); }        // table full (libtiff: CODE_MAX - 1)
      else if (next > (1 << width) - 1) ++width;
      omega = k;
    }
    put(omega);
    // the decoder adds an entry after this code too (and may widen): keep in step before EOI (libtiff LZWPostEncode)
    ++next;
    if (next == 4094) { put(256); reset(); }
    else if (next > (1 << width) - 1) ++width;
  }
  put(257);
  if (nacc > 0) {
    if (out < cap) dst[out] = (unsigned char)(acc << (8 - nacc)); else overflow = true;
    ++out;
  }
  if (overflow) { td_set_error("td_tiff_lzw_encode: output buffer too small"); return TD_ERR_OVERFLOW; }
  return out;
}

// ---- device side: every strip / tile of a raster decoded at once --------------------------------------------
// A raster of 10 000 x 10 000 px has thousands of independent LZW streams.  One warp decodes one stream.  The
// classic decoder walks a prefix chain per code (a pointer chase per output byte); here a table entry is
// {position, length} of an EARLIER OCCURRENCE of its string in the decoded output -- entry `next` is the previous
// string plus the first byte of the current one, and those bytes are contiguous in the output -- so emitting a code
// is a short lane-parallel copy inside the output, and adding an entry is one shared-memory store.  The bit reader
// keeps 64 bits in registers and loads the following 32 one refill ahead.  Same stream rules as the host decoder
// above (MSB-first codes, 9..12 bits, early change, ClearCode 256, EOI 257, a stream may end without EOI).
namespace {

constexpr int kLzwPosBits = 20;                     // chunks of up to 1 MiB decoded bytes
constexpr int kLzwMaxChunk = 1 << kLzwPosBits;

struct LzwBatch {
  const unsigned char* src;     // the file (or any buffer holding the compressed chunks), device memory
  const long long* src_pos;     // (n) first byte of chunk k in src
  const int* src_len;           // (n) compressed bytes of chunk k
  int n;
  unsigned char* dst;           // chunk k is decoded to dst + k * dst_stride
  long long dst_stride;
  const int* dst_len;           // (n) expected decoded bytes of chunk k (<= dst_stride, <= 1 MiB)
  int* out_len;                 // (n) nullable: decoded bytes
  int* status;                  // one int, set to a TD_ERR_* code by the first failing chunk
};

// Decoding one code at a time is one serial dependency chain per warp (~70 instructions per code were measured).
// But between two ClearCodes the WIDTH of every code is a function of its index alone (the table grows by exactly
// one entry per code), so the bit position of the k-th code is known in closed form: the 32 lanes extract 32
// consecutive codes at once.  What remains sequential is tiny: string lengths (a code may name an entry created by
// an earlier code of the same batch: length = that code's predecessor + 1, resolved in rounds over lower lanes), an
// exclusive scan for the output positions, and the copies -- lane-parallel for the usual short strings, all lanes
// together for long ones, in dependency rounds when the source bytes belong to the same batch.
constexpr int kLzwWin = 64;     // window words
constexpr int kLzwLong = 24;    // strings longer than this are copied by the whole warp

TD_D int lzw_bits_before(int k) {        // bits of the first k codes after a ClearCode
  return 9 * k + max(k - 254, 0) + max(k - 766, 0) + max(k - 1790, 0);
}
TD_D int lzw_width(int k) { return 9 + (k >= 254) + (k >= 766) + (k >= 1790); }

__global__ void __launch_bounds__(32) lzw_decode_kernel(LzwBatch a) {
  __shared__ uint32_t tab[4096];                    // length << 20 | position, codes >= 258
  __shared__ uint32_t win[kLzwWin + 1];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x;
  const int c = blockIdx.x;
  if (c >= a.n) return;
  const unsigned char* s = a.src + a.src_pos[c];
  const int n_src = a.src_len[c];
  unsigned char* d = a.dst + (long long)c * a.dst_stride;
  const int cap = a.dst_len[c];
  const int mis = (int)((uintptr_t)s & 3);          // the stream starts `mis` bytes into an aligned word
  const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s - mis);
  const int n_words = (mis + n_src + 3) >> 2;
  const int end_bits = 8 * (mis + n_src);
  int wbase = 0;
  auto fill = [&](int base) {                       // words [base, base + kLzwWin] of the stream, big endian
    __syncwarp();
    for (int k = lane; k <= kLzwWin; k += 32) {
      const int wi = base + k;
      win[k] = wi < n_words ? __byte_perm(__ldg(s4 + wi), 0, 0x0123) : 0u;
    }
    __syncwarp();
  };
  fill(0);
  int base_bit = 8 * mis;       // bit position of code 0 of the current run (a run = the codes between two clears)
  int k = 0;                    // codes of the run consumed so far
  int ppos = -1, plen = 0;      // string of the last consumed code (ppos < 0: none yet in this run)
  int cur = 0, err = 0;
  bool done = false;
  while (!done) {
    // ---- 32 codes ---------------------------------------------------------------------------------------------
    const int idx = k + lane;
    const int bit = base_bit + lzw_bits_before(idx);
    const int width = lzw_width(idx);
    const int first_word = (base_bit + lzw_bits_before(k)) >> 5;
    if (first_word + 14 - wbase > kLzwWin) { wbase = first_word; fill(wbase); }
    const bool inside = bit + width <= end_bits;
    int code = 257;
    if (inside) {
      const int wi = (bit >> 5) - wbase;
      code = (int)(__funnelshift_l(win[wi + 1], win[wi], bit & 31) >> (32 - width));
    }
    const unsigned stops = __ballot_sync(full, code == 256 || code == 257);
    const int nvalid = stops ? __ffs(stops) - 1 : 32;
    const bool live = lane < nvalid;
    // ---- lengths and sources -----------------------------------------------------------------------------------
    // an entry e >= 258 was created by the code of index e - 257 of this run
    int len = 0, spos = -1, dep = -1;          // dep: lane of this batch whose string (+ next first byte) is the source
    bool known = true;
    if (live) {
      if (code < 256) {
        len = 1;
      } else {
        const int creator = code - 257;        // >= 1
        if (creator < k) {
          const uint32_t e = tab[code];
          spos = (int)(e & (kLzwMaxChunk - 1));
          len = (int)(e >> kLzwPosBits);
        } else if (creator <= idx && creator - 1 <= 3838 && (idx > 0)) {
          dep = creator - k - 1;               // -1: the last string of the previous batch
          known = false;
        } else {
          err = TD_ERR_ARG;                    // a code beyond the table
        }
      }
      if (idx == 0 && code >= 256) err = TD_ERR_ARG;     // a table code right after ClearCode
    }
    if (__any_sync(full, err != 0)) { err = TD_ERR_ARG; break; }
    // rounds: a lane learns its length once its source lane knows its own; `round` = copy order
    int round = 0;
    for (int r = 1; __any_sync(full, !known); ++r) {
      const int src = dep < 0 ? 0 : dep;
      const int l_src = __shfl_sync(full, len, src);
      const bool k_src = __shfl_sync(full, (int)known, src) != 0;
      // the entry's string also takes the first byte of lane dep + 1 (for dep + 1 == lane: its own first byte)
      const int nxt = min(dep + 1, 31);
      const bool k_nxt = __shfl_sync(full, (int)known, nxt) != 0;
      if (!known) {
        if (dep < 0) {
          if (dep + 1 == lane || k_nxt) { len = plen + 1; known = true; round = r; }
        } else if (k_src && (dep + 1 == lane || k_nxt)) {
          len = l_src + 1; known = true; round = r;
        }
      }
      if (r > 40) { err = TD_ERR_ARG; break; }
    }
    if (err) break;
    // ---- positions ---------------------------------------------------------------------------------------------
    int incl = len;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += v;
    }
    const int pos = cur + incl - len;
    const int total = __shfl_sync(full, incl, 31);
    if (cur + total > cap) { err = TD_ERR_OVERFLOW; break; }
    const int pos_prev = __shfl_up_sync(full, pos, 1);
    const int len_prev = __shfl_up_sync(full, len, 1);
    const int p_pos = lane == 0 ? ppos : pos_prev, p_len = lane == 0 ? plen : len_prev;
    const int pos_dep = __shfl_sync(full, pos, max(dep, 0));
    if (live && code >= 258 && spos < 0) spos = dep >= 0 ? pos_dep : ppos;
    // ---- table: the code of index idx >= 1 creates entry 257 + idx = previous string + this string's first byte -----
    if (live && idx >= 1 && 257 + idx < 4096) tab[257 + idx] = ((uint32_t)(p_len + 1) << kLzwPosBits) | (uint32_t)p_pos;
    // ---- copies ------------------------------------------------------------------------------------------------
    const int max_round = __reduce_max_sync(full, live ? round : 0);
    for (int r = 0; r <= max_round; ++r) {
      const bool mine = live && round == r;
      if (mine && len <= kLzwLong) {
        if (code < 256) {
          d[pos] = (unsigned char)code;
        } else {
          for (int t = 0; t < len; ++t) d[pos + t] = d[spos + t];     // forward: KwKwK re-reads its own first byte
        }
      }
      unsigned longs = __ballot_sync(full, mine && len > kLzwLong);
      __syncwarp();
      while (longs) {
        const int l = __ffs(longs) - 1;
        longs &= longs - 1;
        const int lp = __shfl_sync(full, pos, l), ls = __shfl_sync(full, spos, l), ll = __shfl_sync(full, len, l);
        // source [ls, ls + ll) ends at or before lp + 1: only the KwKwK last byte overlaps the destination
        const int body = (ls + ll > lp) ? ll - 1 : ll;
        for (int t = lane; t < body; t += 32) d[lp + t] = d[ls + t];
        __syncwarp();
        if (body < ll && lane == 0) d[lp + body] = d[ls + body];
        __syncwarp();
      }
    }
    __syncwarp();
    // ---- next batch ----------------------------------------------------------------------------------------------
    if (nvalid > 0) {
      ppos = __shfl_sync(full, pos, nvalid - 1);
      plen = __shfl_sync(full, len, nvalid - 1);
    }
    cur += total;
    k += nvalid;
    if (nvalid < 32) {
      const int stop_code = __shfl_sync(full, code, nvalid);
      const bool stop_inside = __shfl_sync(full, (int)inside, nvalid) != 0;
      if (stop_code == 256 && stop_inside) {         // ClearCode: a new run starts behind it
        base_bit = __shfl_sync(full, bit + width, nvalid);
        k = 0;
        ppos = -1;
        plen = 0;
      } else {
        done = true;                                  // EOI, or the stream ended without one
      }
    }
  }
  __syncwarp();
  for (int t = cur + lane; t < cap; t += 32) d[t] = 0;   // a short stream leaves zeros (deterministic)
  if (lane == 0) {
    if (a.out_len) a.out_len[c] = cur;
    if (err) atomicCAS(a.status, 0, err);
  }
}

struct PlaceArgs {
  const unsigned char* dec;     // decoded chunks, chunk k at dec + k * stride
  long long stride;
  unsigned char* out;           // (C, H, W) planar raster of `ssize`-byte samples
  int C, H, W, ssize;
  int planar;                   // TIFF PlanarConfiguration: 1 = chunky (samples interleaved), 2 = one plane per chunk
  int chunk_rows, chunk_cols, nx, ny;
  int predictor;                // 1 none, 2 horizontal differencing (8-bit samples)
};

// one warp per chunk row: undo the predictor (a per-channel running sum along the row, modulo 256) and scatter
// the samples into the planar raster
__global__ void __launch_bounds__(128) place_chunks_kernel(PlaceArgs a, int n_chunks) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x;
  if (k >= n_chunks) return;
  const int per_plane = a.nx * a.ny;
  const int plane = a.planar == 2 ? k / per_plane : 0;
  const int rem = a.planar == 2 ? k % per_plane : k;
  const int j = rem / a.nx, i = rem % a.nx;
  const int spp = a.planar == 2 ? 1 : a.C;
  const int y0 = j * a.chunk_rows, x0 = i * a.chunk_cols;
  const int ncols = min(a.chunk_cols, a.W - x0);
  const size_t plane_px = (size_t)a.H * a.W;
  for (int r = blockIdx.y * 4 + (threadIdx.x >> 5); r < a.chunk_rows; r += gridDim.y * 4) {
    const int y = y0 + r;
    if (y >= a.H) break;
    const unsigned char* row = a.dec + (size_t)k * a.stride + (size_t)r * a.chunk_cols * spp * a.ssize;
    if (a.ssize == 4) {        // 32-bit samples, no predictor
      const uint32_t* row4 = reinterpret_cast<const uint32_t*>(row);
      uint32_t* out4 = reinterpret_cast<uint32_t*>(a.out);
      for (int t = lane; t < ncols * spp; t += 32) {
        const int x = t / spp, ch = t - x * spp;
        out4[(size_t)(plane + ch) * plane_px + (size_t)y * a.W + x0 + x] = row4[t];
      }
      continue;
    }
    if (a.predictor != 2) {
      for (int t = lane; t < ncols * spp; t += 32) {
        const int x = t / spp, ch = t - x * spp;
        a.out[(size_t)(plane + ch) * plane_px + (size_t)y * a.W + x0 + x] = row[t];
      }
      continue;
    }
    // predictor 2: lane l owns columns [l * seg, (l + 1) * seg)
    const int seg = (ncols + 31) / 32;
    const int xa = min(lane * seg, ncols), xb = min(xa + seg, ncols);
    uint32_t sum[4] = {0, 0, 0, 0};
    for (int x = xa; x < xb; ++x)
      for (int ch = 0; ch < spp; ++ch) sum[ch] += row[(size_t)x * spp + ch];
    uint32_t off[4];
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v = sum[ch];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      off[ch] = v - sum[ch];                        // exclusive
    }
    for (int x = xa; x < xb; ++x)
      for (int ch = 0; ch < spp; ++ch) {
        off[ch] += row[(size_t)x * spp + ch];
        a.out[(size_t)(plane + ch) * plane_px + (size_t)y * a.W + x0 + x] = (unsigned char)off[ch];
      }
  }
}

}  // namespace

extern "C" int td_tiff_lzw_decode_batch(const unsigned char* src, const long long* src_pos, const int* src_len,
                                        int n_chunks, unsigned char* dst, long long dst_stride, const int* dst_len,
                                        int* out_len, int* status, void* stream) {
  TD_ARG(n_chunks >= 0);
  if (n_chunks == 0) return TD_OK;
  TD_ARG(src && src_pos && src_len && dst && dst_len && status && dst_stride > 0 && dst_stride <= kLzwMaxChunk);
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  LzwBatch a{src, src_pos, src_len, n_chunks, dst, dst_stride, dst_len, out_len, status};
  lzw_decode_kernel<<<n_chunks, 32, 0, st>>>(a);
  TD_CHECK_LAUNCH("td_tiff_lzw_decode_batch");
  return TD_OK;
}

extern "C" int td_tiff_place_chunks(const unsigned char* decoded, long long stride, int n_chunks, void* out, int bands,
                                    int height, int width, int sample_size, int planar, int chunk_rows, int chunk_cols,
                                    int predictor, void* stream) {
  TD_ARG(n_chunks >= 0);
  if (n_chunks == 0) return TD_OK;
  TD_ARG(decoded && out && bands > 0 && height > 0 && width > 0 && chunk_rows > 0 && chunk_cols > 0 && stride > 0);
  TD_ARG(sample_size == 1 || sample_size == 4);
  TD_ARG(planar == 1 || planar == 2);
  TD_ARG(predictor == 1 || (predictor == 2 && sample_size == 1));
  TD_ARG(planar == 2 || bands <= 4);
  const int nx = td_div_up(width, chunk_cols), ny = td_div_up(height, chunk_rows);
  TD_ARG(n_chunks == nx * ny * (planar == 2 ? bands : 1));
  PlaceArgs a{decoded, stride, (unsigned char*)out, bands, height, width, sample_size, planar, chunk_rows, chunk_cols,
              nx, ny, predictor};
  const int gy = chunk_rows >= 64 ? 8 : (chunk_rows >= 8 ? 2 : 1);
  place_chunks_kernel<<<dim3(n_chunks, gy), 128, 0, (cudaStream_t)stream>>>(a, n_chunks);
  TD_CHECK_LAUNCH("td_tiff_place_chunks");
  return TD_OK;
}
