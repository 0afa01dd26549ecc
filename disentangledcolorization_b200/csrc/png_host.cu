// Host-side PNG writer for the CLI's writer threads (no device work; part of the C ABI so that the call releases the GIL).
//
// Replaces `Image.fromarray(rgb).save(path, "PNG")` of the reference's save path (utils/util.py:91-106 as used by
// main/colorizer/inference.py:131-135).  The stored PIXELS are the same; the file bytes are not (PNG leaves filter and
// deflate choices to the encoder): every row uses the Sub filter, the zlib stream is one dynamic-Huffman deflate block of
// literals plus distance-1 runs (flat areas), no LZ77 search.  With the forward on the GPU, PNG encoding was the largest
// host cost of the CLI (OpenCV/libpng level 1 = 7.7 ms per 256 x 256 image, PIL level 6 = 18-26 ms); this
// encoder takes ~0.5 ms for files ~10 % larger than libpng's level 1.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../../include/disco_b200.h"

namespace {

// ---- checksums ----------------------------------------------------------------------------------------------------------
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
const CrcTables g_crc;

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {   // slice-by-8, little-endian host
  crc = ~crc;
  while (n >= 8) {
    uint32_t a, b;
    memcpy(&a, p, 4);
    memcpy(&b, p + 4, 4);
    a ^= crc;
    crc = g_crc.t[7][a & 0xFF] ^ g_crc.t[6][(a >> 8) & 0xFF] ^ g_crc.t[5][(a >> 16) & 0xFF] ^ g_crc.t[4][a >> 24] ^
          g_crc.t[3][b & 0xFF] ^ g_crc.t[2][(b >> 8) & 0xFF] ^ g_crc.t[1][(b >> 16) & 0xFF] ^ g_crc.t[0][b >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = g_crc.t[0][(crc ^ *p++) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

uint32_t adler32(const uint8_t* p, size_t n) {
  uint32_t a = 1, b = 0;
  while (n) {
    size_t k = n < 5552 ? n : 5552;                      // largest run for which b cannot overflow 32 bits
    n -= k;
    while (k--) { a += *p++; b += a; }
    a %= 65521u;
    b %= 65521u;
  }
  return (b << 16) | a;
}

// ---- length-limited canonical Huffman codes --------------------------------------------------------------------------------
// Plain two-queue Huffman construction; if the deepest leaf exceeds `limit`, the frequencies are flattened (halved, kept
// >= 1) and the tree is rebuilt (terminates: equal frequencies give depth ceil(log2 n) <= 9 for n <= 286).
void huffman_lengths(const uint32_t* freq_in, int n, int limit, uint8_t* len) {
  std::vector<uint32_t> freq(freq_in, freq_in + n);
  for (;;) {
    struct Node { uint64_t w; int l, r; };
    std::vector<Node> nodes;
    std::vector<int> leaves;
    for (int i = 0; i < n; ++i) { len[i] = 0; if (freq[i]) leaves.push_back(i); }
    if (leaves.empty()) return;
    if (leaves.size() == 1) { len[leaves[0]] = 1; return; }
    std::stable_sort(leaves.begin(), leaves.end(), [&](int a, int b) { return freq[a] < freq[b]; });
    const int L = (int)leaves.size();
    nodes.reserve(2 * L);
    for (int i = 0; i < L; ++i) nodes.push_back({freq[leaves[i]], -1, leaves[i]});
    int qa = 0, qb = L;                                  // leaf queue [qa, L), internal queue [qb, nodes.size())
    while ((L - qa) + ((int)nodes.size() - qb) > 1) {
      int pick[2];
      for (int k = 0; k < 2; ++k) {
        const bool has_a = qa < L, has_b = qb < (int)nodes.size();
        if (has_a && (!has_b || nodes[qa].w <= nodes[qb].w)) pick[k] = qa++;
        else pick[k] = qb++;
      }
      nodes.push_back({nodes[pick[0]].w + nodes[pick[1]].w, pick[0], pick[1]});
    }
    // depths, root = last node
    std::vector<int> depth(nodes.size(), 0);
    int maxd = 0;
    for (int i = (int)nodes.size() - 1; i >= L; --i) {
      depth[nodes[i].l] = depth[i] + 1;
      depth[nodes[i].r] = depth[i] + 1;
    }
    for (int i = 0; i < L; ++i) { len[nodes[i].r] = (uint8_t)depth[i]; maxd = std::max(maxd, depth[i]); }
    if (maxd <= limit) return;
    for (int i = 0; i < n; ++i) if (freq[i]) freq[i] = (freq[i] + 1) / 2;
    bool all_one = true;
    for (int i = 0; i < n; ++i) if (freq[i] > 1) all_one = false;
    if (all_one) {                                       // cannot happen for n <= 2^limit; guard against an endless loop
      for (int i = 0; i < n; ++i) if (freq[i]) len[i] = (uint8_t)limit;
      return;
    }
  }
}

// canonical code of each symbol, bit-reversed for deflate's LSB-first packing
void canonical_codes(const uint8_t* len, int n, uint16_t* code) {
  int count[16] = {0}, next[16] = {0};
  for (int i = 0; i < n; ++i) count[len[i]]++;
  count[0] = 0;
  int c = 0;
  for (int b = 1; b < 16; ++b) { c = (c + count[b - 1]) << 1; next[b] = c; }
  for (int i = 0; i < n; ++i) {
    const int l = len[i];
    if (!l) { code[i] = 0; continue; }
    unsigned v = next[l]++, r = 0;
    for (int k = 0; k < l; ++k) { r = (r << 1) | (v & 1); v >>= 1; }
    code[i] = (uint16_t)r;
  }
}

struct BitWriter {
  uint8_t* p;
  uint8_t* end;
  uint64_t acc = 0;
  int nbits = 0;
  bool overflow = false;
  inline void put(uint32_t v, int n) {                   // n <= 32
    acc |= (uint64_t)v << nbits;
    nbits += n;
    if (nbits >= 32) {
      if (end - p < 4) { overflow = true; nbits -= 32; acc >>= 32; return; }
      const uint32_t w = (uint32_t)acc;
      memcpy(p, &w, 4);
      p += 4;
      acc >>= 32;
      nbits -= 32;
    }
  }
  void finish() {
    while (nbits > 0) {
      if (p >= end) { overflow = true; return; }
      *p++ = (uint8_t)acc;
      acc >>= 8;
      nbits -= 8;
    }
    nbits = 0;
  }
};

// deflate length symbol tables (RFC 1951 3.2.5)
struct LenTab {
  uint16_t sym[259];
  uint8_t extra_bits[259];
  uint16_t extra_val[259];
  LenTab() {
    static const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const int eb[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    for (int l = 3; l <= 258; ++l) {
      int s = 28;
      while (base[s] > l) --s;
      sym[l] = (uint16_t)(257 + s);
      extra_bits[l] = (uint8_t)eb[s];
      extra_val[l] = (uint16_t)(l - base[s]);
    }
  }
};
const LenTab g_len;

inline void be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

constexpr int kMinRun = 4;      // distance-1 matches shorter than this are cheaper as literals

}  // namespace

extern "C" long long disco_host_png_bound(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const long long raw = (long long)H * (1 + 3LL * W);
  return raw * 2 + 4096;        // 15-bit worst-case codes + header + chunk framing
}

extern "C" int disco_host_png_encode(const uint8_t* rgb, int H, int W, long long row_stride, uint8_t* out, long long cap,
                                     long long* out_len) {
  if (!rgb || !out || !out_len || H <= 0 || W <= 0 || row_stride < 3LL * W) return DISCO_ERR_INVALID;
  if (cap < disco_host_png_bound(H, W)) return DISCO_ERR_INVALID;
  const size_t rowb = 1 + 3 * (size_t)W, nraw = rowb * H;
  std::vector<uint8_t> raw(nraw);
  // Sub filter (type 1): byte - byte three to the left (the same channel of the previous pixel)
  for (int y = 0; y < H; ++y) {
    const uint8_t* s = rgb + (size_t)y * row_stride;
    uint8_t* d = raw.data() + (size_t)y * rowb;
    d[0] = 1;
    d[1] = s[0]; d[2] = s[1]; d[3] = s[2];
    for (size_t x = 3; x < 3 * (size_t)W; ++x) d[1 + x] = (uint8_t)(s[x] - s[x - 3]);
  }
  // pass 1: tokens (literal, or run of the previous byte = match at distance 1) and symbol frequencies
  uint32_t freq[286] = {0};
  std::vector<uint16_t> tok(nraw + 1);                   // < 256 literal; 0x8000 | length for a run
  size_t nt = 0, i = 0;
  while (i < nraw) {
    if (i > 0) {
      const uint8_t prev = raw[i - 1];
      size_t r = 0;
      const size_t lim = std::min<size_t>(258, nraw - i);
      while (r < lim && raw[i + r] == prev) ++r;
      if (r >= (size_t)kMinRun) {
        tok[nt++] = (uint16_t)(0x8000 | r);
        freq[g_len.sym[r]]++;
        i += r;
        continue;
      }
    }
    tok[nt++] = raw[i];
    freq[raw[i]]++;
    ++i;
  }
  freq[256] = 1;
  uint8_t ll_len[286], cl_len[19];
  uint16_t ll_code[286], cl_code[19];
  huffman_lengths(freq, 286, 15, ll_len);
  canonical_codes(ll_len, 286, ll_code);
  int hlit = 286;
  while (hlit > 257 && ll_len[hlit - 1] == 0) --hlit;
  // one distance code (distance 1) of length 1: RFC 1951 3.2.7 "if only one distance code is used, it is encoded using one bit"
  uint32_t cl_freq[19] = {0};
  for (int s = 0; s < hlit; ++s) cl_freq[ll_len[s]]++;
  cl_freq[1]++;                                          // the distance code length
  huffman_lengths(cl_freq, 19, 7, cl_len);
  canonical_codes(cl_len, 19, cl_code);
  static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19;
  while (hclen > 4 && cl_len[order[hclen - 1]] == 0) --hclen;

  // ---- file image: signature, IHDR, IDAT (zlib stream), IEND ----
  uint8_t* p = out;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memcpy(p, sig, 8);
  p += 8;
  be32(p, 13);
  memcpy(p + 4, "IHDR", 4);
  be32(p + 8, (uint32_t)W);
  be32(p + 12, (uint32_t)H);
  p[16] = 8; p[17] = 2; p[18] = 0; p[19] = 0; p[20] = 0;  // 8 bits, colour type 2 (RGB), deflate, adaptive filters, no interlace
  be32(p + 21, crc32_update(0, p + 4, 17));
  p += 25;
  uint8_t* idat = p;                                     // length filled in below
  memcpy(p + 4, "IDAT", 4);
  p += 8;
  *p++ = 0x78;                                           // zlib header: deflate, 32 KB window; FLG makes CMF*256+FLG % 31 == 0
  *p++ = 0x01;
  BitWriter bw{p, out + cap - 16};
  bw.put(1, 1);                                          // BFINAL
  bw.put(2, 2);                                          // BTYPE = dynamic Huffman
  bw.put((uint32_t)(hlit - 257), 5);
  bw.put(0, 5);                                          // HDIST - 1 = 0
  bw.put((uint32_t)(hclen - 4), 4);
  for (int k = 0; k < hclen; ++k) bw.put(cl_len[order[k]], 3);
  for (int s = 0; s < hlit; ++s) bw.put(cl_code[ll_len[s]], cl_len[ll_len[s]]);
  bw.put(cl_code[1], cl_len[1]);                         // distance code 0: length 1
  for (size_t k = 0; k < nt; ++k) {
    const uint16_t t = tok[k];
    if (t < 256) {
      bw.put(ll_code[t], ll_len[t]);
    } else {
      const int r = t & 0x7FFF;
      const int s = g_len.sym[r];
      bw.put(ll_code[s], ll_len[s]);
      if (g_len.extra_bits[r]) bw.put(g_len.extra_val[r], g_len.extra_bits[r]);
      bw.put(0, 1);                                      // distance code 0 (= distance 1), no extra bits
    }
  }
  bw.put(ll_code[256], ll_len[256]);                     // end of block
  bw.finish();
  if (bw.overflow) return DISCO_ERR_INVALID;
  p = bw.p;
  be32(p, adler32(raw.data(), nraw));
  p += 4;
  const uint32_t idat_len = (uint32_t)(p - (idat + 8));
  be32(idat, idat_len);
  be32(p, crc32_update(0, idat + 4, idat_len + 4));
  p += 4;
  be32(p, 0);
  memcpy(p + 4, "IEND", 4);
  be32(p + 8, crc32_update(0, p + 4, 4));
  p += 12;
  *out_len = (long long)(p - out);
  return DISCO_OK;
}

extern "C" int disco_host_png_write(const char* path, const uint8_t* rgb, int H, int W, long long row_stride) {
  if (!path) return DISCO_ERR_INVALID;
  const long long cap = disco_host_png_bound(H, W);
  if (cap <= 0) return DISCO_ERR_INVALID;
  std::vector<uint8_t> buf((size_t)cap);
  long long n = 0;
  const int rc = disco_host_png_encode(rgb, H, W, row_stride, buf.data(), cap, &n);
  if (rc != DISCO_OK) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) return DISCO_ERR_INVALID;
  const size_t w = fwrite(buf.data(), 1, (size_t)n, f);
  const int c = fclose(f);
  return (w == (size_t)n && c == 0) ? DISCO_OK : DISCO_ERR_INVALID;
}
