// 2-bit transport encoding of Dna texts (see transport.cu): the format and the host-side packer.
//
// Format: 8 characters -> 2 bytes.  Byte 2g holds bit 1 of characters 8g .. 8g+7 (character j in
// bit j), byte 2g + 1 holds bit 2: the two bit planes of the Dna code (c >> 1) & 3 (reference
// src/profiles/dna.rs:19-23).  Bit planes are what one GF2P8AFFINEQB produces from 8 text bytes
// (the text bytes are the 8 x 8 bit matrix, a constant selects the rows), so the packer is a load,
// an affine transform and a share of a permute + full-line streaming store per 64 characters; the
// device expands the planes with a few shifts and one PRMT per 4 characters.
//
// Header-only so that the emulation library (g++, no CUDA) can test every host code path
// against the device-side decoder.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define SB_PACK_X86 1
#endif

#include "profile.h"

namespace sb {

// ---- decoder (device + host reference) ------------------------------------------------------
// planes = byte 2g | byte 2g+1 << 8 (one little-endian u16); returns characters 8g + 4 half .. + 3
// as canonical upper-case bytes.  Code -> byte through PRMT on the table "ACTG" (A=0,C=1,T=2,G=3).
SB_HD uint32_t dna_unpack4(uint32_t planes, int half) {
  const uint32_t x = ((planes >> (4 * half)) & 0xFu) | (((planes >> (8 + 4 * half)) & 0xFu) << 16);
  const uint32_t y = (x | (x << 6)) & 0x03030303u;
  const uint32_t z = (y | (y << 3)) & 0x11111111u;  // plane bits at 0,4,8,12 (bit 1) and 16,20,24,28 (bit 2)
  const uint32_t sel = (z | (z >> 15)) & 0x3333u;   // selector nibble of character i = its code
  const uint32_t table = 0x47544341u;               // 'A','C','T','G'
#if defined(__CUDA_ARCH__)
  return __byte_perm(table, 0u, sel);
#else
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= ((table >> (8 * ((sel >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
  return r;
#endif
}

// ---- packer (host) --------------------------------------------------------------------------
// All variants write ceil(n / 8) * 2 bytes and return false if a byte outside ACGTacgt was seen
// (the output is still written).

inline bool dna_pack_scalar(const uint8_t* src, uint8_t* dst, size_t n) {
  bool ok = true;
  for (size_t g = 0; g * 8 < n; g++) {
    uint32_t lo = 0, hi = 0;
    for (size_t j = 0; j < 8 && g * 8 + j < n; j++) {
      const uint8_t c = src[g * 8 + j], u = c & 0xDF;
      ok &= (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
      lo |= (uint32_t)((c >> 1) & 1) << j;
      hi |= (uint32_t)((c >> 2) & 1) << j;
    }
    dst[2 * g] = (uint8_t)lo;
    dst[2 * g + 1] = (uint8_t)hi;
  }
  return ok;
}

#if defined(SB_PACK_X86)
// 32 characters per iteration: the sign-bit gather of two shifted copies gives the planes.
__attribute__((target("avx2"))) inline bool dna_pack_avx2(const uint8_t* src, uint8_t* dst, size_t n) {
  const __m256i up = _mm256_set1_epi8((char)0xDF);
  const __m256i lut = _mm256_broadcastsi128_si256(
      _mm_setr_epi8((char)0xFF, 0x41, 0, 0x43, 0x54, 0, 0, 0x47, 0, 0, 0, 0, 0, 0, 0, 0));  // low nibble -> the one valid byte (0xFF / 0: none)
  __m256i bad = _mm256_setzero_si256();
  size_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i u = _mm256_and_si256(v, up);
    // a byte with bit 7 set selects 0 from the table and differs from itself
    bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_shuffle_epi8(lut, u), u));
    const uint32_t lo = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
    const uint32_t hi = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
    const __m128i il = _mm_unpacklo_epi8(_mm_cvtsi32_si128((int)lo), _mm_cvtsi32_si128((int)hi));
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + (i >> 2)), il);
  }
  bool ok = _mm256_testz_si256(bad, bad) != 0;
  if (i < n) ok &= dna_pack_scalar(src + i, dst + (i >> 2), n - i);
  return ok;
}

// 256 characters per iteration.  GF2P8AFFINEQB with the TEXT as the matrix operand: result byte i
// of a 64-bit lane, bit j = parity(text byte 7 - j & selector byte i), i.e. with one-hot selectors
// 0x02 / 0x04 bytes 0 and 1 of every lane are the two planes of its 8 characters (bit-reversed).
// Two word permutes gather the 32 plane pairs of four registers, a second affine transform
// reverses the bits of every byte, one streaming store writes the full line.  The source is
// prefetched into L2 16 KB ahead: a single core's hardware prefetcher sustains
// less than half of that rate on the hosts this runs on.
inline size_t dna_pack_ahead() {
  static const size_t ahead = [] {
    const char* e = getenv("SASSY_B200_PACK_AHEAD");  // bytes, tuning knob
    return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)16384;
  }();
  return ahead;
}
// `stream`: full-line non-temporal stores (a staging buffer that is written once and read by the copy
// engine from memory); false keeps the lines in the cache (a small staging ring that is re-used).
__attribute__((target("avx512f,avx512bw,gfni"))) inline bool dna_pack_gfni(const uint8_t* src, uint8_t* dst,
                                                                            size_t n, bool stream = true) {
  const __m512i up = _mm512_set1_epi8((char)0xDF);
  const __m512i lut =
      _mm512_broadcast_i32x4(_mm_setr_epi8((char)0xFF, 0x41, 0, 0x43, 0x54, 0, 0, 0x47, 0, 0, 0, 0, 0, 0, 0, 0));
  const __m512i sel = _mm512_set1_epi64(0x0000000000000402ll);
  const __m512i bitrev = _mm512_set1_epi64((long long)0x8040201008040201ull);
  alignas(64) static const uint16_t idx[32] = {0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60};
  const __m512i gather = _mm512_load_si512(idx);
  __m512i bad = _mm512_setzero_si512();
  size_t i = 0;
  const bool aligned = stream && (reinterpret_cast<uintptr_t>(dst) & 63) == 0;
  const size_t kDnaPackAhead = dna_pack_ahead();
  for (; i + 256 <= n; i += 256) {
    for (int q = 0; q < 4; q++) _mm_prefetch(reinterpret_cast<const char*>(src + i + kDnaPackAhead + 64 * q), _MM_HINT_T1);
    const __m512i v0 = _mm512_loadu_si512(src + i), v1 = _mm512_loadu_si512(src + i + 64);
    const __m512i v2 = _mm512_loadu_si512(src + i + 128), v3 = _mm512_loadu_si512(src + i + 192);
    const __m512i u0 = _mm512_and_si512(v0, up), u1 = _mm512_and_si512(v1, up);
    const __m512i u2 = _mm512_and_si512(v2, up), u3 = _mm512_and_si512(v3, up);
    bad = _mm512_ternarylogic_epi64(bad, _mm512_shuffle_epi8(lut, u0), u0, 0xF6);  // bad | (a ^ b)
    bad = _mm512_ternarylogic_epi64(bad, _mm512_shuffle_epi8(lut, u1), u1, 0xF6);
    bad = _mm512_ternarylogic_epi64(bad, _mm512_shuffle_epi8(lut, u2), u2, 0xF6);
    bad = _mm512_ternarylogic_epi64(bad, _mm512_shuffle_epi8(lut, u3), u3, 0xF6);
    const __m512i p0 = _mm512_gf2p8affine_epi64_epi8(sel, v0, 0), p1 = _mm512_gf2p8affine_epi64_epi8(sel, v1, 0);
    const __m512i p2 = _mm512_gf2p8affine_epi64_epi8(sel, v2, 0), p3 = _mm512_gf2p8affine_epi64_epi8(sel, v3, 0);
    const __m512i a = _mm512_permutex2var_epi16(p0, gather, p1);  // low 256 bits: word 0 of every lane
    const __m512i b = _mm512_permutex2var_epi16(p2, gather, p3);
    const __m512i r = _mm512_gf2p8affine_epi64_epi8(_mm512_inserti64x4(a, _mm512_castsi512_si256(b), 1), bitrev, 0);
    if (aligned)
      _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + (i >> 2)), r);  // no read-for-ownership
    else
      _mm512_storeu_si512(dst + (i >> 2), r);
  }
  _mm_sfence();
  bool ok = _mm512_test_epi64_mask(bad, bad) == 0;
  if (i < n) ok &= dna_pack_scalar(src + i, dst + (i >> 2), n - i);
  return ok;
}
#endif  // SB_PACK_X86

// level: 0 scalar, 1 AVX2, 2 AVX-512 + GFNI; -1 = the best the CPU has
inline int dna_pack_best_level() {
#if defined(SB_PACK_X86)
  if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("gfni")) return 2;
  if (__builtin_cpu_supports("avx2")) return 1;
#endif
  return 0;
}

inline bool dna_pack(const uint8_t* src, uint8_t* dst, size_t n, int level = -1, bool stream = true) {
  static const int best = dna_pack_best_level();
  if (level < 0 || level > best) level = best;
#if defined(SB_PACK_X86)
  if (level == 2) return dna_pack_gfni(src, dst, n, stream);
  if (level == 1) return dna_pack_avx2(src, dst, n);
#endif
  return dna_pack_scalar(src, dst, n);
}

}  // namespace sb
