// LZ4 block codec for the proof wire format, host code.
//
// The reference serialises a proof with postcard and compresses it with lz4_flex::compress_prepend_size
// (crates/rec_aggregation/src/type_1_aggregation.rs:81-89): a 4-byte little-endian uncompressed length followed by ONE
// LZ4 block.  lz4_flex 0.13.0 (Cargo.lock) is a crates.io dependency that is not part of the reference tree; this file
// restates the published LZ4 block format (lz4.org "LZ4 Block Format Description"): a block is a sequence of
//   token (high nibble: literal length, low nibble: match length - 4), [literal length extension bytes 255.. < 255],
//   literals, 2-byte little-endian match offset, [match length extension bytes]
// whose last sequence stops after its literals; the last 5 bytes of the input are literals and no match starts within the
// last 12 bytes.  Any block that obeys the format is decoded by lz4_flex::decompress_size_prepended, whichever
// compressor produced it, so the compressor below is a plain greedy one (4-byte hash, 64 K-entry table).
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/leanmultisig_b200.h"

int lm_internal_fail(int code, const char* msg);

namespace {
inline uint32_t read32(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
inline void put_len(std::vector<uint8_t>& out, size_t len) {  // the part of a length beyond the token nibble's 15
  while (len >= 255) {
    out.push_back(255);
    len -= 255;
  }
  out.push_back((uint8_t)len);
}
}  // namespace

extern "C" {

uint64_t lm_lz4_compress_bound(uint64_t n) { return 4 + n + n / 255 + 16; }

int lm_lz4_compress_prepend_size(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_len) {
  if ((!src && n) || !dst || !out_len) return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_compress_prepend_size: null argument");
  if (n > 0xffffffffull) return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_compress_prepend_size: input longer than 2^32 - 1 bytes");
  std::vector<uint8_t> out;
  out.reserve(n / 2 + 64);
  const uint32_t n32 = (uint32_t)n;
  for (int k = 0; k < 4; k++) out.push_back((uint8_t)(n32 >> (8 * k)));
  constexpr int HASH_BITS = 16;
  std::vector<int64_t> table((size_t)1 << HASH_BITS, -1);
  uint64_t anchor = 0, i = 0;
  const uint64_t match_limit = n >= 12 ? n - 12 : 0;  // no match may start beyond this
  while (n >= 13 && i < match_limit) {
    const uint32_t seq = read32(src + i);
    const uint32_t h = (seq * 2654435761u) >> (32 - HASH_BITS);
    const int64_t cand = table[h];
    table[h] = (int64_t)i;
    if (cand >= 0 && i - (uint64_t)cand <= 65535 && read32(src + cand) == seq) {
      uint64_t len = 4;
      const uint64_t max_len = n - 5 - i;  // the last five bytes stay literals
      while (len < max_len && src[cand + len] == src[i + len]) len++;
      const uint64_t lit = i - anchor;
      out.push_back((uint8_t)(((lit < 15 ? lit : 15) << 4) | (len - 4 < 15 ? len - 4 : 15)));
      if (lit >= 15) put_len(out, lit - 15);
      out.insert(out.end(), src + anchor, src + i);
      const uint16_t off = (uint16_t)(i - (uint64_t)cand);
      out.push_back((uint8_t)off);
      out.push_back((uint8_t)(off >> 8));
      if (len - 4 >= 15) put_len(out, len - 4 - 15);
      i += len;
      anchor = i;
    } else {
      i++;
    }
  }
  const uint64_t lit = n - anchor;
  out.push_back((uint8_t)((lit < 15 ? lit : 15) << 4));
  if (lit >= 15) put_len(out, lit - 15);
  out.insert(out.end(), src + anchor, src + n);
  if (out.size() > cap) return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_compress_prepend_size: output buffer too small");
  memcpy(dst, out.data(), out.size());
  *out_len = out.size();
  return LM_OK;
}

int lm_lz4_decompress_size_prepended(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_len) {
  if (!src || !out_len || n < 4) return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_decompress_size_prepended: bad argument");
  const uint64_t want = read32(src);
  *out_len = want;
  if (!dst) return LM_OK;  // size query
  if (want > cap) return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_decompress_size_prepended: output buffer too small");
  uint64_t ip = 4, op = 0;
  auto bad = [] { return lm_internal_fail(LM_ERR_INVALID, "lm_lz4_decompress_size_prepended: malformed block"); };
  if (want == 0) return (n == 5 && src[4] == 0) || n == 4 ? LM_OK : bad();
  while (ip < n) {
    const uint8_t token = src[ip++];
    uint64_t lit = token >> 4;
    if (lit == 15) {
      uint8_t b;
      do {
        if (ip >= n) return bad();
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (ip + lit > n || op + lit > want) return bad();
    memcpy(dst + op, src + ip, lit);
    ip += lit, op += lit;
    if (ip == n) break;  // the last sequence has no match
    if (ip + 2 > n) return bad();
    const uint64_t off = (uint64_t)src[ip] | ((uint64_t)src[ip + 1] << 8);
    ip += 2;
    uint64_t len = (token & 15) + 4;
    if ((token & 15) == 15) {
      uint8_t b;
      do {
        if (ip >= n) return bad();
        b = src[ip++];
        len += b;
      } while (b == 255);
    }
    if (off == 0 || off > op || op + len > want) return bad();
    for (uint64_t k = 0; k < len; k++) dst[op + k] = dst[op - off + k];  // byte by byte: matches may overlap their output
    op += len;
  }
  return op == want ? LM_OK : bad();
}

}  // extern "C"
