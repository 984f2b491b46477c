// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/f128.h header).
// Portable BLAKE3 (hash mode only), restated from the published BLAKE3 specification
// (reference dependency `blake3 1.8.2`, /root/reference/Cargo.lock:110-113; used by
// zk-lisp-proof-winterfell/src/poseidon/mod.rs:421-440 to derive Poseidon constants).
// Pinned by the official known-answer vectors in tests/test_oracle_blake3.py.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

namespace zkl {

struct Blake3 {
  static constexpr uint32_t CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8;
  static const uint32_t *IV() {
    static const uint32_t iv[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A,
                                   0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
    return iv;
  }
  static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  static inline void g(uint32_t *s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
  }
  static void compress(const uint32_t cv[8], const uint8_t block[64], uint64_t counter,
                       uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    static const int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
    uint32_t m[16];
    for (int i = 0; i < 16; i++)
      m[i] = (uint32_t)block[4 * i] | ((uint32_t)block[4 * i + 1] << 8) |
             ((uint32_t)block[4 * i + 2] << 16) | ((uint32_t)block[4 * i + 3] << 24);
    uint32_t s[16];
    for (int i = 0; i < 8; i++) s[i] = cv[i];
    for (int i = 0; i < 4; i++) s[8 + i] = IV()[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = block_len; s[15] = flags;
    for (int r = 0; r < 7; r++) {
      g(s, 0, 4, 8, 12, m[0], m[1]);   g(s, 1, 5, 9, 13, m[2], m[3]);
      g(s, 2, 6, 10, 14, m[4], m[5]);  g(s, 3, 7, 11, 15, m[6], m[7]);
      g(s, 0, 5, 10, 15, m[8], m[9]);  g(s, 1, 6, 11, 12, m[10], m[11]);
      g(s, 2, 7, 8, 13, m[12], m[13]); g(s, 3, 4, 9, 14, m[14], m[15]);
      if (r < 6) { uint32_t t[16]; for (int i = 0; i < 16; i++) t[i] = m[PERM[i]]; memcpy(m, t, sizeof m); }
    }
    for (int i = 0; i < 8; i++) { out[i] = s[i] ^ s[i + 8]; out[i + 8] = s[i + 8] ^ cv[i]; }
  }

  // one-shot hash of `len` bytes -> 32-byte digest
  static void hash(const uint8_t *data, size_t len, uint8_t out[32]) {
    // Split into 1024-byte chunks; keep a CV stack merged per the binary-counter rule.
    std::vector<uint32_t> stack;  // 8 words per entry
    size_t nchunks = len == 0 ? 1 : (len + 1023) / 1024;
    uint32_t root_words[16];
    for (size_t ci = 0; ci < nchunks; ci++) {
      const uint8_t *cp = data + ci * 1024;
      size_t clen = (ci + 1 == nchunks) ? len - ci * 1024 : 1024;
      size_t nblocks = clen == 0 ? 1 : (clen + 63) / 64;
      uint32_t cv[8]; memcpy(cv, IV(), 32);
      bool last_chunk = (ci + 1 == nchunks);
      uint32_t words[16];
      for (size_t bi = 0; bi < nblocks; bi++) {
        uint8_t block[64]; memset(block, 0, 64);
        size_t blen = (bi + 1 == nblocks) ? clen - bi * 64 : 64;
        if (blen) memcpy(block, cp + bi * 64, blen);
        uint32_t flags = 0;
        if (bi == 0) flags |= CHUNK_START;
        if (bi + 1 == nblocks) flags |= CHUNK_END;
        if (bi + 1 == nblocks && last_chunk && nchunks == 1) flags |= ROOT;
        compress(cv, block, ci, (uint32_t)blen, flags, words);
        memcpy(cv, words, 32);
      }
      if (nchunks == 1) { memcpy(root_words, words, sizeof words); break; }
      if (!last_chunk) {
        // push and merge while the number of completed chunks has trailing zero bits
        uint32_t cur[8]; memcpy(cur, cv, 32);
        size_t total = ci + 1;
        while ((total & 1) == 0) {
          uint8_t block[64];
          memcpy(block, &stack[stack.size() - 8], 32); memcpy(block + 32, cur, 32);
          stack.resize(stack.size() - 8);
          uint32_t w[16]; compress(IV(), block, 0, 64, PARENT, w); memcpy(cur, w, 32);
          total >>= 1;
        }
        stack.insert(stack.end(), cur, cur + 8);
      } else {
        // final chunk: fold the stack right-to-left; last merge is ROOT
        uint32_t cur[8]; memcpy(cur, cv, 32);
        while (!stack.empty()) {
          uint8_t block[64];
          memcpy(block, &stack[stack.size() - 8], 32); memcpy(block + 32, cur, 32);
          stack.resize(stack.size() - 8);
          uint32_t flags = PARENT | (stack.empty() ? ROOT : 0);
          uint32_t w[16]; compress(IV(), block, 0, 64, flags, w);
          memcpy(cur, w, 32);
          if (stack.empty()) memcpy(root_words, w, sizeof w);
        }
      }
    }
    for (int i = 0; i < 8; i++) {
      out[4 * i] = (uint8_t)root_words[i]; out[4 * i + 1] = (uint8_t)(root_words[i] >> 8);
      out[4 * i + 2] = (uint8_t)(root_words[i] >> 16); out[4 * i + 3] = (uint8_t)(root_words[i] >> 24);
    }
  }
};

}  // namespace zkl
