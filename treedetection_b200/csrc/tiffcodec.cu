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
